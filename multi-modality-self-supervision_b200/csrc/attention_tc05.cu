// attention_tc05.cu — fused masked attention for sm_100a, forward and backward, head dim 64, bf16 operands.
//   * Q/K/V tiles arrive by TMA straight out of the packed [B*L, 3H] QKV activation (no head-major transpose);
//   * S = QK^T, PV, and the five backward products run on tcgen05.mma with fp32 accumulators in TMEM;
//   * the MedViLL image/text block mask is evaluated from (mode, A, t_len) per element (mask.cuh) — no mask tensor —
//     and KV / Q tiles that are fully masked for the sample's mode are skipped (exact: the reference's additive
//     -10000 underflows to 0 after softmax, models/cxrbert_origin.py:82-83);
//   * probabilities never touch HBM: forward keeps one row per thread (online softmax), backward recomputes P from
//     the saved row log-sum-exp; attention-probability dropout is regenerated from a counter-based RNG.
// Reference arithmetic: upstream BertSelfAttention (twin: .../pytorch_pretrained_bert/model.py:301-320).
#include <limits.h>

#include "attn_common.cuh"
#include "kernels.h"
#include "tc05.cuh"
#include "tmap.h"

// Phase timeline (debug build `make tests/attn_timeline`, -DMV_ATTN_TIMELINE): two probe threads of one mid-grid CTA
// record clock64() at the marks; compiled out otherwise.
#ifdef MV_ATTN_TIMELINE
#define TL_DECL(slot, cond)                                                                  \
  unsigned long long* tl_p = (a.timeline && (cond)) ? a.timeline + (slot) * 64 : nullptr; \
  int tl_n = 0
#define TL_MARK() do { if (tl_p && tl_n < 64) tl_p[tl_n++] = clock64(); } while (0)
#else
#define TL_DECL(slot, cond) do { } while (0)
#define TL_MARK() do { } while (0)
#endif

namespace mv {
using namespace tc05;

namespace {

constexpr int D = 64;
constexpr int TQ = 128, TK = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr uint32_t TILE_BYTES = TQ * D * 2;       // 16 KB: one [128 x 64] bf16 tile
constexpr uint32_t P_BYTES = TQ * TK * 2;         // 32 KB: [128 x 128] bf16 as two 64-column halves

struct FwdSmem {
  uint64_t bar_q, bar_k, bar_v, bar_s, bar_o;
  uint32_t tmem_base;
  float xmax[2][2][TQ];    // [tile parity][column half][row]
  float xsum[2][TQ];
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// P row `r`: 32 consecutive columns starting at c32*32, from 16 already-packed bf16x2 words, into the swizzled
// [128 x 128] tile (two 64-column halves)
__device__ __forceinline__ void store_pk32(uint8_t* sP, int r, int c32, const uint32_t (&pk)[16]) {
  uint8_t* half = sP + (c32 >> 1) * TILE_BYTES;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(half + sw128_off(r, (c32 & 1) * 4 + j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

// Forward.  256 threads: two threads per query row (warp w reads TMEM lane quarter w & 3 and owns the column half
// w >> 2 of S and of O), two CTAs per SM.  Online softmax with a reference exponent that lags one key tile behind the
// running maximum (exact after the final normalisation); the reference is seeded from the first 16 keys.
__global__ void __launch_bounds__(256, 2)
attn_fwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = smem + 2 * TILE_BYTES;
  uint8_t* sP = smem + 3 * TILE_BYTES;
  FwdSmem* sh = reinterpret_cast<FwdSmem*>(smem + 3 * TILE_BYTES + P_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, ch = warp >> 2;
  const int r = lq * 32 + lane;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int L = a.L, H = a.nh * D, A = a.A;
  const int mode = a.mode[b], tl = a.t_len[b];
  const int q_lo = qt * TQ, q_hi = min(q_lo + TQ - 1, L - 1);
  const int n_kv = (L + TK - 1) / TK;
  const int row0 = b * L;  // first row of this sample in the [B*L, 3H] matrix

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&sh->bar_q, 1); mbar_init(&sh->bar_k, 1); mbar_init(&sh->bar_v, 1);
    mbar_init(&sh->bar_s, 1); mbar_init(&sh->bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&sh->tmem_base, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(lq * 32) << 16);

  auto active = [&](int j) { return tile_any_allowed(mode, q_lo, q_hi, j * TK, min(j * TK + TK - 1, L - 1), A, tl); };
  auto next_active = [&](int j) { ++j; while (j < n_kv && !active(j)) ++j; return j; };
  int j = next_active(-1);

  if (tid == 0) {
    mbar_expect_tx(&sh->bar_q, TILE_BYTES);
    tma_load_2d(&tmQKV, &sh->bar_q, sQ, h * D, row0 + q_lo);
    if (j < n_kv) {
      mbar_expect_tx(&sh->bar_k, TILE_BYTES);
      tma_load_2d(&tmQKV, &sh->bar_k, sK, H + h * D, row0 + j * TK);
      mbar_expect_tx(&sh->bar_v, TILE_BYTES);
      tma_load_2d(&tmQKV, &sh->bar_v, sV, 2 * H + h * D, row0 + j * TK);
    }
  }

  constexpr uint32_t idesc_s = make_idesc_bf16(TQ, TK, 0, 0);
  constexpr uint32_t idesc_o = make_idesc_bf16(TQ, D, 0, 1);
  const float scale2 = 0.125f * kLog2e;
  const int q = q_lo + r;
  int m_lo, m_hi;
  mask_row_interval(mode, min(q, L - 1), A, tl, L, m_lo, m_hi);
  const uint32_t m_span = static_cast<uint32_t>(m_hi - m_lo);
  // warp-level bounds of the rows' allowed key intervals (rows past the sequence end excluded): a 32-column chunk starting
  // at c_lo can hold an allowed entry only if c_lo < w_hi and c_lo + 31 >= w_lo; chunks that cannot cost one zero store
  int w_lo = q < L ? m_lo : INT_MAX, w_hi = q < L ? m_hi : INT_MIN;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o));
    w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
  }
  float o_acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
  float ref = 0.f, l_loc = 0.f;
  uint32_t it = 0;

  for (; j < n_kv; ++it) {
    const int jn = next_active(j);
    const uint32_t ph = it & 1u;
    const int k_lo = j * TK;
    if (warp == 0) {      // warp-uniform (descriptors stay in uniform registers); one elected lane issues
      if (it == 0) mbar_wait(&sh->bar_q, 0);
      mbar_wait(&sh->bar_k, ph);
      tc_fence_after();
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16(tmem, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(&sh->bar_s);
      }
      __syncwarp();
    }
    mbar_wait(&sh->bar_s, ph);
    tc_fence_after();
    if (tid == 0 && jn < n_kv) {  // K buffer is free: prefetch the next active key tile under the softmax
      mbar_expect_tx(&sh->bar_k, TILE_BYTES);
      tma_load_2d(&tmQKV, &sh->bar_k, sK, H + h * D, row0 + jn * TK);
    }
    const bool full = tile_all_allowed(mode, q_lo, q_hi, k_lo, k_lo + TK - 1, A, tl) && (k_lo + TK <= L);
    const int kc0 = k_lo + ch * 64;     // first key column of this thread's half
    if (it == 0) {
      // Seed the reference exponent from the first 16 keys of the first tile (both threads of a row read the same
      // columns, so they agree without an exchange).  Any finite reference is exact after the final normalisation —
      // softmax is shift invariant and the running reference catches up with the true maximum at the end of this tile —
      // so the seed only has to be within the fp32 exponent range (2^126) of the row maximum.  A full max pass over the
      // tile cost ~4 000 cycles per CTA: executed once per CTA, its code missed the 32 KB L1.5 instruction cache every
      // time (profiles/r01_attn_timeline.txt).
      uint32_t v[16];
      tmem_ld16(t_lane, v);
      tmem_ld_wait();
      float mx = -INFINITY, mx_any = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float sv = __uint_as_float(v[i]);
        mx_any = fmaxf(mx_any, sv);
        mx = fmaxf(mx, (static_cast<uint32_t>(k_lo + i - m_lo) < m_span) ? sv : -INFINITY);
      }
      ref = (mx == -INFINITY ? mx_any : mx) * scale2;
    }
    float m_raw = -INFINITY;                 // running max of the RAW scores of this tile
    const int n_kk = (min(TK, L - k_lo) + 15) >> 4;     // the PV product spans only the key columns that exist
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int c_lo = kc0 + c * 32;
      if (c_lo - k_lo >= 16 * n_kk) break;               // beyond the PV product's K extent: never read
      if (!(c_lo < w_hi && c_lo + 31 >= w_lo)) {         // no row of this warp sees these keys: P = 0
        const uint32_t zero[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        store_pk32(sP, r, ch * 2 + c, zero);
        continue;
      }
      uint32_t v[32];
      tmem_ld32(t_lane + ch * 64 + c * 32, v);
      tmem_ld_wait();
      float p[32];
      // this row's allowed keys inside the chunk are i in [rel_lo, rel_hi): all / none / a sub-range
      const int rel_lo = m_lo - (kc0 + c * 32), rel_hi = m_hi - (kc0 + c * 32);
      if (full || (rel_lo <= 0 && rel_hi >= 32)) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(v[i]);
          m_raw = fmaxf(m_raw, s);
          p[i] = ex2(fmaf(s, scale2, -ref));        // exp2(s * scale2 - ref): one FFMA + one MUFU
          l_loc += p[i];
        }
      } else {                       // partial or empty (span = 0 gives p = 0 everywhere): one code path, smaller kernel
        const uint32_t span = rel_hi > rel_lo ? static_cast<uint32_t>(rel_hi - rel_lo) : 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (static_cast<uint32_t>(i - rel_lo) < span) ? __uint_as_float(v[i]) : -INFINITY;
          m_raw = fmaxf(m_raw, s);
          p[i] = ex2(fmaf(s, scale2, -ref));
          l_loc += p[i];
        }
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
      if (a.drop_on) {
        // dropped probabilities are zeroed with one AND per bf16 pair; the 1/(1-p) keep-scale is applied once to O
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint64_t grp = attn_drop_group(b, a.nh, h, L, min(q, L - 1), (kc0 + c * 32 + 16 * g) >> 4);
          const uint4 rnd = dropout_rand16(a.drop, a.drop_site, grp);
          const uint32_t t4 = dropout_thresh4(a.drop, grp);
          const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t fl = keep_flags4_any(rw[w], t4);
            pk[8 * g + 2 * w] &= keep_mask_pair(fl, 0);
            pk[8 * g + 2 * w + 1] &= keep_mask_pair(fl, 1);
          }
        }
      }
      store_pk32(sP, r, ch * 2 + c, pk);
    }
    sh->xmax[ph][ch][r] = m_raw * scale2;
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      mbar_wait(&sh->bar_v, ph);
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
      if (elect_one()) {
        for (int kk = 0; kk < n_kk; ++kk)
          umma_bf16(tmem + 128, make_smem_desc_sw128(pa + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(va + kk * 2048, 8192, 1024), idesc_o, kk > 0);
        umma_commit(&sh->bar_o);
      }
      __syncwarp();
    }
    const float m_new = fmaxf(sh->xmax[ph][0][r], sh->xmax[ph][1][r]);
    mbar_wait(&sh->bar_o, ph);
    tc_fence_after();
    if (tid == 0 && jn < n_kv) {  // V buffer is free
      mbar_expect_tx(&sh->bar_v, TILE_BYTES);
      tma_load_2d(&tmQKV, &sh->bar_v, sV, 2 * H + h * D, row0 + jn * TK);
    }
    {
      uint32_t v[32];
      tmem_ld32(t_lane + 128 + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] += __uint_as_float(v[i]);
    }
    if (m_new > ref) {  // move the reference exponent up (both threads of a row take the same decision)
      const float alpha = ex2(ref - m_new);
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] *= alpha;
      l_loc *= alpha;
      ref = m_new;
    }
    j = jn;
  }

  sh->xsum[ch][r] = l_loc;
  tc_fence_before();
  __syncthreads();
  if (q < L) {
    const float l_row = sh->xsum[0][r] + sh->xsum[1][r];
    const float inv = l_row > 0.f ? (a.drop_on ? a.drop.scale : 1.f) / l_row : 0.f;   // deferred dropout keep-scale
    bf16* dst = static_cast<bf16*>(a.ctx) + (static_cast<long>(row0) + q) * H + h * D + ch * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 u;
      u.x = pack_bf16x2(o_acc[8 * c + 0] * inv, o_acc[8 * c + 1] * inv);
      u.y = pack_bf16x2(o_acc[8 * c + 2] * inv, o_acc[8 * c + 3] * inv);
      u.z = pack_bf16x2(o_acc[8 * c + 4] * inv, o_acc[8 * c + 5] * inv);
      u.w = pack_bf16x2(o_acc[8 * c + 6] * inv, o_acc[8 * c + 7] * inv);
      reinterpret_cast<uint4*>(dst)[c] = u;
    }
    if (ch == 0) a.lse[(static_cast<long>(b) * a.nh + h) * L + q] = ref * kLn2 + logf(l_row);
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// ------------------------------------------------------------------------------------------------------------------
// delta[b,h,q] = sum_d dO[q,d] * O[q,d]   (one warp per row of the [B*L, H] activations)
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ O,
                                                         float* __restrict__ delta, int rows, int L, int nh) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int H = nh * D, b = warp / L, q = warp % L;
  for (int ch = lane; ch < (H >> 3); ch += 32) {
    const uint4 x = *reinterpret_cast<const uint4*>(dO + static_cast<long>(warp) * H + ch * 8);
    const uint4 y = *reinterpret_cast<const uint4*>(O + static_cast<long>(warp) * H + ch * 8);
    const float2 a0 = unpack_bf16x2(x.x), a1 = unpack_bf16x2(x.y), a2 = unpack_bf16x2(x.z), a3 = unpack_bf16x2(x.w);
    const float2 b0 = unpack_bf16x2(y.x), b1 = unpack_bf16x2(y.y), b2 = unpack_bf16x2(y.z), b3 = unpack_bf16x2(y.w);
    float s = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0) delta[(static_cast<long>(b) * nh + (ch >> 3)) * L + q] = s;
  }
}

__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv,
                                                              long rows, int H) {
  pdl_sync();
  const long n8 = rows * (H >> 3);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / (H >> 3);
    const int c = static_cast<int>(i % (H >> 3)) * 8;
    const float4 x = *reinterpret_cast<const float4*>(dq_acc + r * H + c), y = *reinterpret_cast<const float4*>(dq_acc + r * H + c + 4);
    uint4 u;
    u.x = pack_bf16x2(x.x, x.y); u.y = pack_bf16x2(x.z, x.w); u.z = pack_bf16x2(y.x, y.y); u.w = pack_bf16x2(y.z, y.w);
    *reinterpret_cast<uint4*>(dqkv + r * 3 * H + c) = u;
  }
}

// Deterministic mode: dQ[b, q, :] = sum over the key tiles that can see query tile (q / 128), in ASCENDING key-tile order, of
// the per-key-tile partials the backward CTAs stored — the same terms the fp32 reduce-adds would have summed in arrival order.
__global__ void __launch_bounds__(256) attn_dq_sum_convert_kernel(const float* __restrict__ part, bf16* __restrict__ dqkv, AttnArgs a) {
  pdl_sync();
  const int H = a.nh * D, L = a.L, n_kv = (L + TK - 1) / TK;
  const long rows = static_cast<long>(a.B) * L;
  const long n8 = rows * (H >> 3);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / (H >> 3);
    const int c = static_cast<int>(i % (H >> 3)) * 8;
    const int b = static_cast<int>(r / L), q = static_cast<int>(r % L);
    const int mode = a.mode[b], tl = a.t_len[b];
    const int qt = q / TQ, q_lo = qt * TQ, q_hi = min(q_lo + TQ - 1, L - 1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int kt = 0; kt < n_kv; ++kt) {
      if (!tile_any_allowed(mode, q_lo, q_hi, kt * TK, min(kt * TK + TK - 1, L - 1), a.A, tl)) continue;
      const float* src = part + ((static_cast<long>(kt) * a.B + b) * L + q) * H + c;
      const float4 x = *reinterpret_cast<const float4*>(src), y = *reinterpret_cast<const float4*>(src + 4);
      acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w; acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
    }
    uint4 u;
    u.x = pack_bf16x2(acc[0], acc[1]); u.y = pack_bf16x2(acc[2], acc[3]); u.z = pack_bf16x2(acc[4], acc[5]); u.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dqkv + r * 3 * H + c) = u;
  }
}

struct BwdSmem {
  uint64_t bar_kv, bar_q[2], bar_s[2], bar_h0, bar_o, bar_dvk, bar_pd, bar_stage, bar_qf;
  uint32_t tmem_base;
};

// 16 consecutive columns starting at `col` (multiple of 16) of row `r`, from 8 packed bf16x2 words
__device__ __forceinline__ void store_pk16(uint8_t* sP, int r, int col, const uint32_t (&pk)[8]) {
  uint8_t* half = sP + (col >> 6) * TILE_BYTES;
  const int j0 = (col & 63) >> 3;
  *reinterpret_cast<uint4*>(half + sw128_off(r, j0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(half + sw128_off(r, j0 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// Backward.  One CTA per (key tile, head, sample) loops over the query tiles that can see this key tile.
// 544 threads = 16 math warps (four threads per query row, each owning 32 of the 128 key columns) + 1 issuer warp (TMA
// loads, tcgen05.mma, dQ TMA reduce-add).  The phases that used to run back to back inside the CTA — wait for S / dP,
// softmax-backward math, the dV / dK / dQ products, dQ staging (~7 000 cycles per query tile, of which ~2 600 are math:
// profiles/r01_attn_timeline.txt) — are software pipelined across query tiles:
//   * P / dS staging is double buffered, so the products of tile i (which read it) run under the math of tile i+1;
//   * S / dP of tile i+1 are issued BEFORE the products of tile i, in two 64-key-column halves: every math warp works
//     through its share of half 0 first and signals (bar_h0) once it holds those scores in registers, so half 0 of the NEXT
//     tile is issued in the middle of this tile's math and is complete when the warps come around (r01 issued the whole
//     next S / dP only after the math phase: 300 - 1 700 cycles of S wait per tile);
//   * dQ of tile i is drained from TMEM at the end of tile i+1's math and staged in the (by then dead) P buffer of
//     tile i; the issuer sends it off as one fp32 TMA reduce-add per 32-column half.
// TMEM columns: S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ [384,448)
// mbarriers (completion k belongs to the k-th ACTIVE query tile of this CTA, parity k & 1):
//   bar_s[h]   half h (key columns [64 h, 64 h + 64)) of S / dP of tile k complete (tcgen05.commit)   math warps wait
//   bar_h0     every math thread has loaded its half-0 scores of tile k from TMEM    512 arrivals, issuer waits
//   bar_pd     P / dS of tile k stored and dQ of tile k-1 staged     512 arrivals, issuer waits (one extra for the tail)
//   bar_dvk    LAST tile only: its dV / dK products (issued after dQ, which then owns bar_o) complete    math warps wait
//   bar_qf     dV / dK products of tile k (not the last) complete: its Q / dO buffer may be refilled    issuer waits
//   bar_o      dV / dK / dQ products of tile k complete              math warps wait before draining dQ(k); issuer before
//                                                                    re-using the Q / dO buffer
//   bar_stage  the reduce-add of dQ(k) has read its staging          issuer arrives, math warps wait before tile k+2's stores
__global__ void __launch_bounds__(544, 1)
attn_bwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDQP,
                     const __grid_constant__ CUtensorMap tmDKV, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;     // [2]
  uint8_t* sdO = smem + 4 * TILE_BYTES;    // [2]
  uint8_t* sPD = smem + 6 * TILE_BYTES;    // [2] x { P 32 KB | dS 32 KB }; the P half doubles as the fp32 dQ staging of the same tile
  BwdSmem* sh = reinterpret_cast<BwdSmem*>(sPD + 4 * P_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int L = a.L, H = a.nh * D, A = a.A;
  const int k_lo = kt * TK, k_hi = min(k_lo + TK - 1, L - 1);
  const int n_q = (L + TQ - 1) / TQ;
  const int row0 = b * L;
  TL_DECL(tid == 512 ? 2 : 3, kt == 1 && h == 3 && b == a.B / 2 && (tid == 512 || tid == 96));
  TL_MARK();

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDQP);
    tma_prefetch_desc(&tmDKV);
    mbar_init(&sh->bar_dvk, 1);
    mbar_init(&sh->bar_qf, 1);
    mbar_init(&sh->bar_kv, 1); mbar_init(&sh->bar_q[0], 1); mbar_init(&sh->bar_q[1], 1);
    mbar_init(&sh->bar_s[0], 1); mbar_init(&sh->bar_s[1], 1); mbar_init(&sh->bar_h0, 512);
    mbar_init(&sh->bar_o, 1); mbar_init(&sh->bar_pd, 512); mbar_init(&sh->bar_stage, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&sh->tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  pdl_sync();
  const int mode = a.mode[b], tl = a.t_len[b];

  auto active = [&](int i) { return tile_any_allowed(mode, i * TQ, min(i * TQ + TQ - 1, L - 1), k_lo, k_hi, A, tl); };
  auto next_active = [&](int i) { ++i; while (i < n_q && !active(i)) ++i; return i; };
  int i = next_active(-1);
  const bool any = i < n_q;

  constexpr uint32_t idesc_s = make_idesc_bf16(TQ, TK / 2, 0, 0); // S, dP : K-major x K-major, one 64-key half per issue
  constexpr uint32_t idesc_t = make_idesc_bf16(TK, D, 1, 1);      // dV, dK: P^T / dS^T (MN-major) x dO / Q (MN-major)
  constexpr uint32_t idesc_q = make_idesc_bf16(TQ, D, 0, 1);      // dQ    : dS (K-major) x K (MN-major)

  if (warp == 16) {
    // ------------------------------ issuer warp: TMA, tcgen05.mma, dQ reduce-add (one elected lane) ------------------------------
    auto load_q = [&](int qi, int buf) {
      mbar_expect_tx(&sh->bar_q[buf], 2 * TILE_BYTES);
      tma_load_2d(&tmQKV, &sh->bar_q[buf], sQ + buf * TILE_BYTES, h * D, row0 + qi * TQ);
      tma_load_2d(&tmDO, &sh->bar_q[buf], sdO + buf * TILE_BYTES, h * D, row0 + qi * TQ);
    };
    // half hf of S = Q K^T and dP = dO V^T for the query tile in buffer `buf`: key rows [64 hf, 64 hf + 64) of K / V
    auto issue_s_dp = [&](uint32_t buf, uint32_t hf) {
      const uint32_t qa = smem_u32(sQ + buf * TILE_BYTES), ka = smem_u32(sK) + hf * (TILE_BYTES / 2),
                     va = smem_u32(sV) + hf * (TILE_BYTES / 2), da = smem_u32(sdO + buf * TILE_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16(tmem + 64 * hf, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16(tmem + 128 + 64 * hf, make_smem_desc_sw128(da + k * 32, 16, 1024), make_smem_desc_sw128(va + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(&sh->bar_s[hf]);
      }
      __syncwarp();
    };
    int in = any ? next_active(i) : n_q;
    if (any) {
      if (lane == 0) {
        mbar_expect_tx(&sh->bar_kv, 2 * TILE_BYTES);
        tma_load_2d(&tmQKV, &sh->bar_kv, sK, H + h * D, row0 + k_lo);
        tma_load_2d(&tmQKV, &sh->bar_kv, sV, 2 * H + h * D, row0 + k_lo);
        load_q(i, 0);
        if (in < n_q) load_q(in, 1);
      }
      __syncwarp();
      mbar_wait(&sh->bar_kv, 0);
      mbar_wait(&sh->bar_q[0], 0);
      tc_fence_after();
      issue_s_dp(0, 0);
      issue_s_dp(0, 1);
    }
    int prev = -1;
    uint32_t it = 0;
    for (; i < n_q; ++it) {
      const uint32_t buf = it & 1u;
      const int inn = in < n_q ? next_active(in) : n_q;
      TL_MARK();
      if (in < n_q) {                             // mid-math of this tile: its half-0 scores are in registers -> next tile's half 0
        mbar_wait(&sh->bar_q[buf ^ 1u], ((it + 1) >> 1) & 1u);
        mbar_wait(&sh->bar_h0, it & 1u);
        tc_fence_after();
        issue_s_dp(buf ^ 1u, 0);
      }
      mbar_wait(&sh->bar_pd, it & 1u);            // P / dS of this tile stored; dQ of the previous tile staged
      tc_fence_after();
      TL_MARK();
      if (it > 0 && lane == 0) {                  // previous tile's dQ: one fp32 reduce-add per 32-column half
        uint8_t* stg = sPD + (buf ^ 1u) * 2 * P_BYTES;
        if (a.dq_part) {                           // deterministic: plain store into this key tile's partial slot (rows >= L clipped)
          tma_store_3d(&tmDQP, stg, h * D, prev * TQ, kt * a.B + b);
          tma_store_3d(&tmDQP, stg + TILE_BYTES, h * D + 32, prev * TQ, kt * a.B + b);
        } else {
          tma_reduce_add_2d(&tmDQ, stg, h * D, row0 + prev * TQ);
          tma_reduce_add_2d(&tmDQ, stg + TILE_BYTES, h * D + 32, row0 + prev * TQ);
        }
        tma_commit_group();
      }
      __syncwarp();
      if (in < n_q) issue_s_dp(buf ^ 1u, 1);      // next tile's second half before this tile's products
      if (it > 0) {                               // the reduce-add has read its staging: tile it+1 may overwrite that buffer.
        // Before the products: their 24 tcgen05.mma take ~2 000 cycles to issue (queue back-pressure), and the math warps of
        // tile it+1 need this buffer ~1 500 cycles into their phase (ncu: 9 % of the kernel's samples sat in that wait)
        if (lane == 0) { tma_wait_group_read<0>(); mbar_arrive(&sh->bar_stage); }
        __syncwarp();
      }
      {
        const uint32_t pa = smem_u32(sPD + buf * 2 * P_BYTES), sa = pa + P_BYTES, qa = smem_u32(sQ + buf * TILE_BYTES),
                       ka = smem_u32(sK), da = smem_u32(sdO + buf * TILE_BYTES);
        const bool last = in >= n_q;              // last tile: dQ first (its drain + reduce-add run under the dV / dK products)
        if (elect_one()) {
          if (last) {
#pragma unroll
            for (int kk = 0; kk < TK / 16; ++kk)  // dQ[q,d] = sum_k dS[q,k] K[k,d]
              umma_bf16(tmem + 384, make_smem_desc_sw128(sa + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 16, 1024),
                        make_smem_desc_sw128(ka + kk * 2048, 8192, 1024), idesc_q, kk > 0);
            umma_commit(&sh->bar_o);
          }
#pragma unroll
          for (int kk = 0; kk < TQ / 16; ++kk)  // dV[k,d] += sum_q P[q,k] dO[q,d]
            umma_bf16(tmem + 256, make_smem_desc_sw128(pa + kk * 2048, TILE_BYTES, 1024),
                      make_smem_desc_sw128(da + kk * 2048, 8192, 1024), idesc_t, (it > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < TQ / 16; ++kk)  // dK[k,d] += sum_q dS[q,k] Q[q,d]
            umma_bf16(tmem + 320, make_smem_desc_sw128(sa + kk * 2048, TILE_BYTES, 1024),
                      make_smem_desc_sw128(qa + kk * 2048, 8192, 1024), idesc_t, (it > 0 || kk > 0));
          if (last) {
            umma_commit(&sh->bar_dvk);
          } else {
            umma_commit(&sh->bar_qf);             // Q / dO of this tile are free once dV / dK are done: dQ does not read them
#pragma unroll
            for (int kk = 0; kk < TK / 16; ++kk)
              umma_bf16(tmem + 384, make_smem_desc_sw128(sa + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 16, 1024),
                        make_smem_desc_sw128(ka + kk * 2048, 8192, 1024), idesc_q, kk > 0);
            umma_commit(&sh->bar_o);
          }
        }
        __syncwarp();
      }
      TL_MARK();
      if (inn < n_q) {                            // Q / dO two tiles ahead: this buffer is free once dV / dK are complete
        mbar_wait(&sh->bar_qf, it & 1u);
        if (lane == 0) load_q(inn, static_cast<int>(buf));
        __syncwarp();
      }
      prev = i;
      i = in;
      in = inn;
    }
    if (any) {                                    // tail: the last tile's dQ (staged in the OTHER buffer: the last tile's own
      mbar_wait(&sh->bar_pd, it & 1u);            // P / dS are still being read by its dV / dK products)
      if (lane == 0) {
        uint8_t* stg = sPD + (it & 1u) * 2 * P_BYTES;
        if (a.dq_part) {
          tma_store_3d(&tmDQP, stg, h * D, prev * TQ, kt * a.B + b);
          tma_store_3d(&tmDQP, stg + TILE_BYTES, h * D + 32, prev * TQ, kt * a.B + b);
        } else {
          tma_reduce_add_2d(&tmDQ, stg, h * D, row0 + prev * TQ);
          tma_reduce_add_2d(&tmDQ, stg + TILE_BYTES, h * D + 32, row0 + prev * TQ);
        }
        tma_commit_group();
        // only the staging buffer has to outlive the copy: wait until the bulk operation has READ shared memory, not until its
        // global reduction has completed (~1-2 us; the grid's completion orders it before the next kernel as for every
        // TMA-store epilogue)
        tma_wait_group_read<0>();
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    return;
  }

  // ------------------------------------------------ math warps (512 threads) ------------------------------------------------
  const int lq = warp & 3, cq = warp >> 2;          // TMEM lane quarter, column quarter
  const int r = lq * 32 + lane;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(lq * 32) << 16);
  const float scale2 = 0.125f * kLog2e;
  auto row_stats = [&](int qi, float& lse2_o, float& delta_o) {
    const int qq = qi * TQ + r;
    const long st = (static_cast<long>(b) * a.nh + h) * L + (qq < L ? qq : 0);
    lse2_o = qq < L ? a.lse[st] : 0.f;               // natural-log LSE; the caller scales by log2(e) where it is consumed, so
    delta_o = qq < L ? a.delta[st] : 0.f;            // that nothing depends on these loads until the next tile
  };
  auto drain_dq = [&](uint32_t buf) {
    // dQ tile (x 1/8) -> swizzled fp32 staging in the dead P buffer `buf`.  Rows past the sequence end hold exact zeros.
    uint32_t v[16];
    tmem_ld16(t_lane + 384 + cq * 16, v);
    tmem_ld_wait();
    uint8_t* half = sPD + buf * 2 * P_BYTES + (cq >> 1) * TILE_BYTES;
#pragma unroll
    for (int e = 0; e < 16; e += 4)
      *reinterpret_cast<float4*>(half + sw128_off(r, (cq & 1) * 4 + (e >> 2))) =
          make_float4(0.125f * __uint_as_float(v[e]), 0.125f * __uint_as_float(v[e + 1]), 0.125f * __uint_as_float(v[e + 2]),
                      0.125f * __uint_as_float(v[e + 3]));
  };
  float lse_raw = 0.f, delta = 0.f;
  if (any) row_stats(i, lse_raw, delta);
  uint32_t it = 0;

  for (; i < n_q; ++it) {
    const int in = next_active(i);
    const uint32_t buf = it & 1u;
    uint8_t* sP = sPD + buf * 2 * P_BYTES;
    uint8_t* sdS = sP + P_BYTES;
    const int q_lo = i * TQ;
    const int q = q_lo + r;
    const bool q_ok = q < L;
    float lse_n = 0.f, delta_n = 0.f;
    if (in < n_q) row_stats(in, lse_n, delta_n);         // next tile's row statistics: latency hidden behind this tile
    const float lse2 = lse_raw * kLog2e;
    int m_lo, m_hi;
    mask_row_interval(mode, q_ok ? q : 0, A, tl, L, m_lo, m_hi);
    const bool full = tile_all_allowed(mode, q_lo, min(q_lo + TQ - 1, L - 1), k_lo, k_lo + TK - 1, A, tl) && (k_lo + TK <= L) &&
                      (q_lo + TQ <= L);
    // warp-level bounds of the rows' allowed key intervals (rows past the sequence end excluded): a 16-column chunk that no
    // row of this warp can see is stored as zeros without touching TMEM, the exponentials or the dropout generator
    int w_lo = q_ok ? m_lo : INT_MAX, w_hi = q_ok ? m_hi : INT_MIN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o));
      w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
    }
    TL_MARK();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col = c * 64 + cq * 16;                  // column inside the 128-wide tile: every warp takes half 0 first
      uint32_t pk[8], dk[8];
      mbar_wait(&sh->bar_s[c], it & 1u);
      tc_fence_after();
      if (c == 0) TL_MARK();
      const bool seen = k_lo + col < w_hi && k_lo + col + 15 >= w_lo;
      uint32_t sv[16], dv[16];
      if (seen) {
        tmem_ld16(t_lane + col, sv);
        tmem_ld16(t_lane + 128 + col, dv);
        tmem_ld_wait();
      }
      if (c == 0) {                                      // half 0 of S / dP may be overwritten by the next tile's
        tc_fence_before();
        mbar_arrive(&sh->bar_h0);
      }
      if (seen) {
      // Work saved per element: the softmax scale 1/8 of dS is folded into the dQ / dK epilogues, the dropout
      // keep-scale of P into the dV epilogue, and dropped entries are zeroed with integer ANDs on the keep bytes.
      float p[16];
      const int rel_lo = m_lo - (k_lo + col), rel_hi = m_hi - (k_lo + col);
      if (full || (q_ok && rel_lo <= 0 && rel_hi >= 16)) {
#pragma unroll
        for (int e = 0; e < 16; ++e) p[e] = ex2(fmaf(__uint_as_float(sv[e]), scale2, -lse2));
      } else if (!q_ok || rel_hi <= 0 || rel_lo >= 16) {
#pragma unroll
        for (int e = 0; e < 16; ++e) p[e] = 0.f;
      } else {
        const uint32_t span = static_cast<uint32_t>(rel_hi - rel_lo);
#pragma unroll
        for (int e = 0; e < 16; ++e)
          p[e] = (static_cast<uint32_t>(e - rel_lo) < span) ? ex2(fmaf(__uint_as_float(sv[e]), scale2, -lse2)) : 0.f;
      }
      if (a.drop_on) {
        const uint64_t grp = attn_drop_group(b, a.nh, h, L, q_ok ? q : 0, (k_lo + col) >> 4);
        const uint4 rnd = dropout_rand16(a.drop, a.drop_site, grp);
        const uint32_t t4 = dropout_thresh4(a.drop, grp);
        const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t fl = keep_flags4_any(rw[w], t4);
          // per-lane masks: byte j of `fl` has bit 7 set iff element 4 w + j is kept
          const uint32_t m01 = keep_mask_pair(fl, 0), m23 = keep_mask_pair(fl, 1);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int e = 4 * w + 2 * hh;
            const uint32_t mp = hh ? m23 : m01;
            // dP masked by the keep masks (all-ones / zero per element), then dS = p * (dP * keep_scale - delta)
            const float t0 = __uint_as_float(dv[e] & keep_mask_elem(fl, 2 * hh));
            const float t1 = __uint_as_float(dv[e + 1] & keep_mask_elem(fl, 2 * hh + 1));
            dk[e >> 1] = pack_bf16x2(p[e] * fmaf(t0, a.drop.scale, -delta), p[e + 1] * fmaf(t1, a.drop.scale, -delta));
            pk[e >> 1] = pack_bf16x2(p[e], p[e + 1]) & mp;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          dk[e >> 1] = pack_bf16x2(p[e] * (__uint_as_float(dv[e]) - delta), p[e + 1] * (__uint_as_float(dv[e + 1]) - delta));
          pk[e >> 1] = pack_bf16x2(p[e], p[e + 1]);
        }
      }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { pk[e] = 0u; dk[e] = 0u; }
      }
      // this staging buffer last held tile it-2: its products were waited for when dQ(it-2) was drained; the
      // reduce-add of that dQ (staged in the P half) must have read it too
      if (c == 0 && it >= 2) mbar_wait(&sh->bar_stage, (it - 2) & 1u);
      store_pk16(sP, r, col, pk);
      store_pk16(sdS, r, col, dk);
    }
    TL_MARK();
    if (it > 0) {      // previous tile's products are complete (they ran under this tile's math): drain its dQ
      mbar_wait(&sh->bar_o, (it - 1) & 1u);
      tc_fence_after();
      TL_MARK();
      drain_dq(buf ^ 1u);
    }
    tc_fence_before();
    fence_proxy_async_smem();
    mbar_arrive(&sh->bar_pd);
    TL_MARK();
    i = in;
    lse_raw = lse_n;
    delta = delta_n;
  }
  if (any) {           // tail: the last tile's dQ (its products were issued first), then dV / dK
    mbar_wait(&sh->bar_o, (it - 1) & 1u);
    tc_fence_after();
    TL_MARK();
    // staged in the buffer the last tile does NOT use (its own P / dS are still being read by the dV / dK products); that
    // buffer's previous occupant, dQ of tile it-2, must have been read by its reduce-add
    if (it >= 2) mbar_wait(&sh->bar_stage, (it - 2) & 1u);
    drain_dq(it & 1u);
    tc_fence_before();
    fence_proxy_async_smem();
    mbar_arrive(&sh->bar_pd);
    TL_MARK();
    mbar_wait(&sh->bar_dvk, 0);
    tc_fence_after();
    TL_MARK();
  }

  // epilogue: dV, dK rows of this key tile -> bf16 -> swizzled staging (the dead dS half of the buffer the last tile does not
  // use) -> one TMA store each.  The 3-D map [B][L][3H] clips key rows past the sequence end.  (r01 / early r02 stored
  // straight from registers: 32 B per thread at a 4.6 KB row stride = one L1 transaction per lane, 4 500 cycles per CTA.)
  // `any` is uniform across the CTA (an unseen key tile gets exact zeros).
  uint8_t* stg_kv = sPD + (it & 1u) * 2 * P_BYTES + P_BYTES;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    uint32_t v[16];
    if (any) {
      tmem_ld16(t_lane + 256 + which * 64 + cq * 16, v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = 0u;
    }
    // deferred factors: dropout keep-scale for dV (which == 0), softmax scale 1/8 for dK
    const float fs = which == 0 ? (a.drop_on ? a.drop.scale : 1.f) : 0.125f;
    uint8_t* tile = stg_kv + which * TILE_BYTES;
#pragma unroll
    for (int e = 0; e < 16; e += 8) {
      uint4 u;
      u.x = pack_bf16x2(fs * __uint_as_float(v[e]), fs * __uint_as_float(v[e + 1]));
      u.y = pack_bf16x2(fs * __uint_as_float(v[e + 2]), fs * __uint_as_float(v[e + 3]));
      u.z = pack_bf16x2(fs * __uint_as_float(v[e + 4]), fs * __uint_as_float(v[e + 5]));
      u.w = pack_bf16x2(fs * __uint_as_float(v[e + 6]), fs * __uint_as_float(v[e + 7]));
      *reinterpret_cast<uint4*>(tile + sw128_off(r, 2 * cq + (e >> 3))) = u;
    }
  }
  fence_proxy_async_smem();
  asm volatile("bar.sync 1, 512;" ::: "memory");             // the 16 math warps (the issuer warp is not part of it)
  if (tid == 0) {
    tma_store_3d(&tmDKV, stg_kv, 2 * H + h * D, k_lo, b);                   // dV
    tma_store_3d(&tmDKV, stg_kv + TILE_BYTES, H + h * D, k_lo, b);          // dK
    tma_commit_group();
    tma_wait_group_read<0>();
  }
  TL_MARK();
  tc_fence_before();
  __syncthreads();
  TL_MARK();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

constexpr uint32_t kFwdSmem = 1024 + 3 * TILE_BYTES + P_BYTES + sizeof(FwdSmem) + 64;
constexpr uint32_t kBwdSmem = 1024 + 6 * TILE_BYTES + 4 * P_BYTES + sizeof(BwdSmem) + 64;

}  // namespace

// r01 forward (one query tile per CTA, serial S -> softmax -> PV); kept for A/B runs and for dropout thresholds > 128
int attention_fwd_legacy_tc05(const AttnArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.qkv && a.ctx && a.lse && a.mode && a.t_len, "attention_fwd: null argument");
  const int H = a.nh * D;
  CUtensorMap tm;
  int rc = tmap_encode_2d(&tm, TMAP_BF16, a.qkv, 3 * H, static_cast<uint64_t>(a.B) * a.L, static_cast<uint64_t>(3 * H) * 2, D, TQ);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    MV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    attr = true;
  }
  dim3 grid((a.L + TQ - 1) / TQ, a.nh, a.B);
  attn_fwd_tc05_kernel<<<grid, 256, kFwdSmem, s>>>(tm, a);
  MV_LAUNCH_CHECK();
  return 0;
}

int attention_bwd_tc05(const AttnArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.qkv && a.ctx && a.lse && a.dctx && a.dqkv && a.dq_acc && a.delta, "attention_bwd: null argument");
  const int H = a.nh * D;
  const long rows = static_cast<long>(a.B) * a.L;
  CUtensorMap tmQKV, tmDO;
  int rc = tmap_encode_2d(&tmQKV, TMAP_BF16, a.qkv, 3 * H, rows, static_cast<uint64_t>(3 * H) * 2, D, TQ);
  if (rc) return rc;
  rc = tmap_encode_2d(&tmDO, TMAP_BF16, a.dctx, H, rows, static_cast<uint64_t>(H) * 2, D, TQ);
  if (rc) return rc;
  CUtensorMap tmDQ;
  rc = tmap_encode_2d(&tmDQ, TMAP_F32, a.dq_acc, H, rows, static_cast<uint64_t>(H) * 4, 32, TQ);
  if (rc) return rc;
  CUtensorMap tmDQP = tmDQ;
  if (a.dq_part) {     // [n_kv * B][L][H] fp32, box 32 x 128 x 1: query rows past a sample's end are clipped instead of spilling over
    const int n_kv = (a.L + TK - 1) / TK;
    rc = tmap_encode_3d(&tmDQP, TMAP_F32, a.dq_part, H, a.L, static_cast<uint64_t>(n_kv) * a.B, static_cast<uint64_t>(H) * 4,
                        static_cast<uint64_t>(a.L) * H * 4, 32, TQ, 1);
    if (rc) return rc;
  }
  CUtensorMap tmDKV;   // dqkv as [B][L][3H] bf16, box 64 x 128 x 1: dV / dK tiles, key rows past a sample's end clipped
  rc = tmap_encode_3d(&tmDKV, TMAP_BF16, a.dqkv, 3 * H, a.L, a.B, static_cast<uint64_t>(3 * H) * 2,
                      static_cast<uint64_t>(a.L) * 3 * H * 2, D, TK, 1);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    MV_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    attr = true;
  }
  if (!a.dq_part) MV_CUDA_CHECK(cudaMemsetAsync(a.dq_acc, 0, rows * H * sizeof(float), s));
  attn_delta_kernel<<<static_cast<int>((rows * 32 + 255) / 256), 256, 0, s>>>(static_cast<const bf16*>(a.dctx),
                                                                             static_cast<const bf16*>(a.ctx), a.delta,
                                                                             static_cast<int>(rows), a.L, a.nh);
  MV_LAUNCH_CHECK();
  dim3 grid((a.L + TK - 1) / TK, a.nh, a.B);
  // (the delta kernel above follows a memset: launched with full stream ordering; it only triggers its dependents early)
  MV_CUDA_CHECK(launch_pdl(attn_bwd_tc05_kernel, grid, dim3(544), kBwdSmem, s, tmQKV, tmDO, tmDQ, tmDQP, tmDKV, a));
  MV_LAUNCH_CHECK();
  if (a.dq_part) MV_CUDA_CHECK(launch_pdl(attn_dq_sum_convert_kernel, dim3(148 * 4), dim3(256), 0, s, static_cast<const float*>(a.dq_part), static_cast<bf16*>(a.dqkv), a));
  else MV_CUDA_CHECK(launch_pdl(attn_dq_convert_kernel, dim3(148 * 4), dim3(256), 0, s, static_cast<const float*>(a.dq_acc), static_cast<bf16*>(a.dqkv), rows, H));
  MV_LAUNCH_CHECK();
  return 0;
}

}  // namespace mv
