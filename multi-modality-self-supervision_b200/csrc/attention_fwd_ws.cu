// attention_fwd_ws.cu — fused masked attention FORWARD for sm_100a, warp-specialised and persistent (r02).
//
// One CTA per SM walks the work items (sample b, head h, PAIR of 128-row query tiles) round-robin.  Roles (640 threads):
//   warps 0-7   softmax group 0 (query tile 2p), warps 8-15 softmax group 1 (query tile 2p + 1): two threads per query row
//               (64 of the 128 key columns each);
//   warp 16     TMA producer: Q pair (double buffered across items), K / V tiles through two-stage rings;
//   warp 17     tcgen05.mma issuer: S_w = Q_w K^T (128 x 128 x 64) and O_w = P_w V (128 x 64 x 128) for both query tiles
//               (the service warps carry the HIGHEST warp ids: the scheduler arbitrates high-id-first, and an issuer that
//               queues behind four busy softmax warps adds its whole instruction stream to every S round trip).
// The two softmax groups ping-pong: while group 0 turns S_0(j) into P_0(j), the tensor pipe computes S_1(j) and
// O_1 += P_1(j-1) V(j-1), and vice versa, so the MUFU / ALU work of one group hides the MMA + barrier latency of the other
// (the r01 kernel ran S -> softmax -> PV serially per CTA: tensor pipe 12 % active, profiles/r01_ncu_attn_fwd_summary.txt).
// S_w(j+1) is issued as soon as P_w(j) has been stored, BEFORE O_w += P_w(j) V(j), so a group never waits for its own PV.
//
// TMEM (512 columns allocated): S_0 [0,128) | S_1 [128,256) | O_0 [256,320) | O_1 [320,384).  O_w accumulates in TMEM over
// the key tiles of an item.  Softmax: p = exp2(s * scale - ref) with a per-row reference exponent `ref` seeded from the first
// 16 keys and moved up LAZILY — only when the running row maximum exceeds it by more than 2^8 (softmax is shift invariant, so
// any reference that keeps p finite is exact after the final normalisation; bf16 / fp32 share the exponent range).  When it
// does move, the row's owner rescales its O row in TMEM (tcgen05.ld / st) after the previous PV product has completed and
// before it releases the next one; at random-init / trained BERT score scales this path is essentially never taken, so the
// common path keeps NO accumulator in registers.
// Mask: closed-form predicate of (mode, A, t_len) per row (mask.cuh); key tiles no row of a query tile can see are skipped
// (no load, no MMA), 32-column chunks no row of a WARP can see cost one zero store, and the PV product only spans the key
// columns that exist (L = 436: the last key tile contributes 64 of 128 columns).
// Dropout on the probabilities: Philox keep bits per (b, h, q, 16 keys) as in the backward kernel and the SIMT twin, applied as
// an AND on packed bf16 pairs with byte-sliced threshold compares (common.cuh: keep_flags4 / keep_mask_pair).
// Reference arithmetic: upstream BertSelfAttention (twin: .../pytorch_pretrained_bert/model.py:301-320).
#include <limits.h>

#include "attn_common.cuh"
#include "kernels.h"
#include "tc05.cuh"
#include "tmap.h"

namespace mv {
using namespace tc05;

namespace {

constexpr int D = 64;
constexpr int TQ = 128, TK = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr uint32_t TILE_BYTES = TQ * D * 2;       // 16 KB: one [128 x 64] bf16 tile
constexpr uint32_t P_BYTES = TQ * TK * 2;         // 32 KB: [128 x 128] bf16 as two 64-column halves
constexpr int kThreads = 640;        // 4 service warps + 2 softmax groups x 8 warps

struct WsBars {
  uint64_t q_full[2], q_empty[2];        // per Q buffer (item parity)
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];   // per ring stage
  uint64_t s_full[2], p_full[2];         // per softmax group
  uint64_t o_full[2];                    // per softmax group: its PV product of the current key tile is complete
  uint32_t tmem_base;
  float xmax[2][2][2][TQ];               // [group][tile parity][column half][row]: scaled row maximum of a key tile
  float xsum[2][2][TQ];                  // [group][column half][row]: row sums at the end of an item
};

constexpr uint32_t OFF_Q = 0;                          // [2 buffers][2 tiles]
constexpr uint32_t OFF_K = 4 * TILE_BYTES;             // [2 stages]
constexpr uint32_t OFF_V = 6 * TILE_BYTES;             // [2 stages]
constexpr uint32_t OFF_P = 8 * TILE_BYTES;             // [2 groups] x 32 KB
constexpr uint32_t OFF_BARS = 8 * TILE_BYTES + 2 * P_BYTES;
constexpr uint32_t kWsSmem = 1024 + OFF_BARS + sizeof(WsBars) + 64;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// P row `r`: 32 consecutive columns starting at c32*32, from 16 packed bf16x2 words, into the swizzled [128 x 128] tile
__device__ __forceinline__ void store_pk32(uint8_t* sP, int r, int c32, const uint32_t (&pk)[16]) {
  uint8_t* half = sP + (c32 >> 1) * TILE_BYTES;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(half + sw128_off(r, (c32 & 1) * 4 + j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

template <int N>
__device__ __forceinline__ void set_maxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void set_maxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

struct Item {
  int b, h, qp, mode, tl;
  uint32_t act0, act1; // bit j: key tile j is visible to some row of query tile 2 qp + w (n_kv <= 32)
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs a, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  WsBars* sh = reinterpret_cast<WsBars*>(smem + OFF_BARS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L, H = a.nh * D, A = a.A;
  const int n_q = (L + TQ - 1) / TQ, n_kv = n_q, n_qp = (n_q + 1) >> 1;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh->q_full[i], 1); mbar_init(&sh->q_empty[i], 2);
      mbar_init(&sh->k_full[i], 1); mbar_init(&sh->k_empty[i], 2);
      mbar_init(&sh->v_full[i], 1); mbar_init(&sh->v_empty[i], 2);
      mbar_init(&sh->s_full[i], 1); mbar_init(&sh->p_full[i], 8);
      mbar_init(&sh->o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 18) { tmem_alloc(&sh->tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  pdl_sync();

  auto decode = [&](int item) {
    Item it;
    it.qp = item % n_qp;
    const int bh = item / n_qp;
    it.h = bh % a.nh;
    it.b = bh / a.nh;
    it.mode = a.mode[it.b];
    it.tl = a.t_len[it.b];
    // key tile j is visible to some row of query tile qt (exact: skipping an invisible tile does not change the result)
    it.act0 = it.act1 = 0u;
#pragma unroll 1
    for (int x = 0; x < 2 * n_kv; ++x) {
      const int w = x & 1, j = x >> 1;
      const int qt = 2 * it.qp + w;
      if (qt < n_q && tile_any_allowed(it.mode, qt * TQ, min(qt * TQ + TQ - 1, L - 1), j * TK, min(j * TK + TK - 1, L - 1), A, it.tl)) {
        if (w) it.act1 |= 1u << j; else it.act0 |= 1u << j;
      }
    }
    return it;
  };
  auto act_mask = [](const Item& it, int w) { return w ? it.act1 : it.act0; };
  auto act = [&](const Item& it, int w, int j) { return ((act_mask(it, w) >> j) & 1u) != 0u; };
  // first key tile after j that query tile w sees (n_kv if none)
  auto next_act = [&](const Item& it, int w, int j) {
    const uint32_t rest = j >= 31 ? 0u : (act_mask(it, w) >> (j + 1)) << (j + 1);
    return rest ? __ffs(rest) - 1 : n_kv;
  };

  // register budget per warpgroup (setmaxnreg is warpgroup-wide): warps 0-3 (producer, issuer, idle) shrink, the two softmax
  // groups (one thread per query row: 64 fp32 accumulators + a 32-column chunk in flight) grow
  // (the setmaxnreg must sit INSIDE its role's branch: ptxas budgets the registers of a region from the setmaxnreg that
  // dominates it; placed before the role dispatch, the softmax code was compiled against the producers' small budget and
  // spilled its running sums to local memory inside the hot loop)
  if (warp >= 16) {
  set_maxnreg_dec<40>();         // pool = 96 regs x 640 threads (launch allocation): 128 x 40 + 512 x 104 <= 61 440
  if (warp == 16) {
    // =============================================== TMA producer ===============================================
    uint32_t kc = 0, ic = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
      const Item it = decode(item);
      const int row0 = it.b * L;
      const uint32_t qb = ic & 1u;
      const bool two = 2 * it.qp + 1 < n_q;
      mbar_wait(&sh->q_empty[qb], ((ic >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&sh->q_full[qb], two ? 2 * TILE_BYTES : TILE_BYTES);
        tma_load_2d(&tmQKV, &sh->q_full[qb], smem + OFF_Q + (qb * 2 + 0) * TILE_BYTES, it.h * D, row0 + 2 * it.qp * TQ);
        if (two) tma_load_2d(&tmQKV, &sh->q_full[qb], smem + OFF_Q + (qb * 2 + 1) * TILE_BYTES, it.h * D, row0 + (2 * it.qp + 1) * TQ);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        if (!(act(it, 0, j) || act(it, 1, j))) continue;
        const uint32_t st = kc & 1u, ph = (kc >> 1) & 1u;
        mbar_wait(&sh->k_empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&sh->k_full[st], TILE_BYTES);
          tma_load_2d(&tmQKV, &sh->k_full[st], smem + OFF_K + st * TILE_BYTES, H + it.h * D, row0 + j * TK);
        }
        __syncwarp();
        mbar_wait(&sh->v_empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&sh->v_full[st], TILE_BYTES);
          tma_load_2d(&tmQKV, &sh->v_full[st], smem + OFF_V + st * TILE_BYTES, 2 * H + it.h * D, row0 + j * TK);
        }
        __syncwarp();
        ++kc;
      }
    }
  } else if (warp == 17) {
    // =============================================== MMA issuer ===============================================
    constexpr uint32_t idesc_s = make_idesc_bf16(TQ, TK, 0, 0);     // S : K-major x K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(TQ, D, 0, 1);      // O : P (K-major) x V (MN-major)
    uint32_t kc0 = 0, ic = 0;
    uint32_t p_cnt[2] = {0u, 0u}, o_cnt[2] = {0u, 0u};
    const uint32_t q_base = smem_u32(smem + OFF_Q), k_base = smem_u32(smem + OFF_K), v_base = smem_u32(smem + OFF_V),
                   p_base = smem_u32(smem + OFF_P);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
      const Item it = decode(item);
      const uint32_t qb = ic & 1u;
      // union rank of key tile j inside this item (ring position = kc0 + rank)
      const uint32_t uni = it.act0 | it.act1;
      auto rank_of = [&](int j) { return __popc(uni & ((1u << j) - 1u)); };
      int js[2], rs[2];
#pragma unroll
      for (int w = 0; w < 2; ++w) { js[w] = next_act(it, w, -1); rs[w] = js[w] < n_kv ? rank_of(js[w]) : 0; }
      const uint32_t o_first[2] = {o_cnt[0], o_cnt[1]};
      mbar_wait(&sh->q_full[qb], (ic >> 1) & 1u);
      tc_fence_after();
      // S_w(js[w]) = Q_w . K(js[w])^T  ->  TMEM columns [128 w, +128)
      auto issue_s = [&](int w) {
        const int j = js[w];
        const uint32_t slot = kc0 + static_cast<uint32_t>(rs[w]);
        const uint32_t st = slot & 1u, ph = (slot >> 1) & 1u;
        mbar_wait(&sh->k_full[st], ph);
        tc_fence_after();
        // descriptors: one 64-bit base per operand tile, advanced by +2 (32 bytes along K inside the 128-byte swizzle row)
        const uint64_t dq = make_smem_desc_sw128(q_base + (qb * 2 + w) * TILE_BYTES, 16, 1024);
        const uint64_t dk = make_smem_desc_sw128(k_base + st * TILE_BYTES, 16, 1024);
        const bool other = act(it, w ^ 1, j);
        const int jn = next_act(it, w, j);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) umma_bf16(tmem + 128 * w, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
          umma_commit(&sh->s_full[w]);
          umma_commit(&sh->k_empty[st]);
          if (!other) mbar_arrive(&sh->k_empty[st]);       // the other query tile does not read this key tile
          if (jn >= n_kv) umma_commit(&sh->q_empty[qb]);   // last S of this query tile: its Q buffer may be refilled
        }
        __syncwarp();
        if (jn < n_kv) rs[w] = rank_of(jn);
        js[w] = jn;
      };
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (js[w] < n_kv) issue_s(w);
        else if (elect_one()) mbar_arrive(&sh->q_empty[qb]);   // query tile absent / sees nothing: release its share of Q
      }
      __syncwarp();
      int rank = 0;
      for (int j = 0; j < n_kv; ++j) {
        const bool a0 = act(it, 0, j), a1 = act(it, 1, j);
        if (!(a0 || a1)) continue;
        const uint32_t slot = kc0 + static_cast<uint32_t>(rank);
        const uint32_t st = slot & 1u, ph = (slot >> 1) & 1u;
        const int n_kk = (min(TK, L - j * TK) + 15) >> 4;        // PV spans only the key columns that exist
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          if (!(w == 0 ? a0 : a1)) continue;
          mbar_wait(&sh->p_full[w], p_cnt[w] & 1u);            // P_w(j) stored, S_w(j) fully read
          ++p_cnt[w];
          tc_fence_after();
          if (js[w] < n_kv) issue_s(w);                        // next S of this group first: ready when it comes around
          mbar_wait(&sh->v_full[st], ph);
          tc_fence_after();
          // P: 16 keys = +2 inside a 64-column half, halves TILE_BYTES (1024 x 16 B) apart; V (MN-major): 16 key rows = +128
          const uint64_t dp = make_smem_desc_sw128(p_base + w * P_BYTES, 16, 1024);
          const uint64_t dv = make_smem_desc_sw128(v_base + st * TILE_BYTES, 8192, 1024);
          const bool other = w == 0 ? a1 : a0;
          const bool acc = o_cnt[w] != o_first[w];             // O_w accumulates over the key tiles of this item
          if (elect_one()) {
            for (int kk = 0; kk < n_kk; ++kk)
              umma_bf16(tmem + 256 + 64 * w, dp + static_cast<uint64_t>((kk >> 2) * (TILE_BYTES >> 4) + (kk & 3) * 2),
                        dv + static_cast<uint64_t>(kk * 128), idesc_o, (acc || kk > 0) ? 1u : 0u);
            umma_commit(&sh->o_full[w]);
            umma_commit(&sh->v_empty[st]);
            if (!other) mbar_arrive(&sh->v_empty[st]);
          }
          __syncwarp();
          ++o_cnt[w];
        }
        ++rank;
      }
      kc0 += static_cast<uint32_t>(rank);
    }
  }
  } else {
    set_maxnreg_inc<104>();
    // =============================================== softmax groups ===============================================
    // 8 warps per group: TMEM lane quarter lq = warp % 4 (32 query rows), column half ch: TWO threads per query row, each
    // owning two of the four 32-column chunks of S / P (interleaved: ch, ch + 2) and 32 of the 64 columns of O.  Four softmax warps per scheduler: a warp
    // that has just issued a MUFU.EX2 (8 issue cycles on the 16-lane XU) or waits on a Philox dependency leaves the slot
    // to three others (one thread per row left two warps per scheduler at 0.27 IPC, profiles/r02_attn_fwd_ws_notes.md).
    const int w = warp >> 3;
    const int w8 = warp & 7;
    const int lq = w8 & 3, ch = w8 >> 2;
    const int r = lq * 32 + lane;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(lq * 32) << 16);
    const uint32_t t_s = t_lane + 128 * w + 32 * ch, t_o = t_lane + 256 + 64 * w + 32 * ch;
    uint8_t* sPw = smem + OFF_P + w * P_BYTES;
    const float scale2 = 0.125f * kLog2e;
    const bool drop_on = a.drop_on != 0;
    constexpr float kLazy = 8.f;                         // move the reference only when the row maximum is > 2^8 above it
    uint32_t s_cnt = 0, o_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode(item);
      const int qt = 2 * it.qp + w;
      if (qt >= n_q) continue;
      const int row0 = it.b * L;
      const int q_lo = qt * TQ;
      const int q = q_lo + r, qc = min(q, L - 1);
      const int qw_lo = q_lo + lq * 32;
      const bool warp_oob = qw_lo >= L;                  // no row of this warp exists: keep the barrier protocol, skip the math
      int m_lo, m_hi;
      mask_row_interval(it.mode, qc, A, it.tl, L, m_lo, m_hi);
      // warp-level bounds of the rows' allowed key intervals (rows past the sequence end excluded): a 32-column chunk
      // [c_lo, c_lo + 31] has SOME allowed entry only if c_lo < hi_max and c_lo + 31 >= lo_min (conservative: a false
      // positive just evaluates the per-row predicate), and is allowed for EVERY row iff lo_max <= c_lo and c_lo + 32 <= hi_min
      int lo_min = q < L ? m_lo : INT_MAX, lo_max = q < L ? m_lo : INT_MIN, hi_min = q < L ? m_hi : INT_MAX, hi_max = q < L ? m_hi : INT_MIN;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo_min = min(lo_min, __shfl_xor_sync(0xffffffffu, lo_min, o));
        lo_max = max(lo_max, __shfl_xor_sync(0xffffffffu, lo_max, o));
        hi_min = min(hi_min, __shfl_xor_sync(0xffffffffu, hi_min, o));
        hi_max = max(hi_max, __shfl_xor_sync(0xffffffffu, hi_max, o));
      }
      const bool warp_whole = qw_lo + 32 <= L;           // all 32 rows of this warp exist
      float ref = 0.f, l_loc = 0.f, alpha_pend = 1.f;
      bool have_prev = false;
      uint32_t tcount = 0;                               // key tiles of this item done so far (parity of the xmax exchange)
      auto wait_prev_o = [&]() {                          // the PV product of the previous key tile has completed
        mbar_wait(&sh->o_full[w], (o_cnt - 1u) & 1u);
        tc_fence_after();
      };
      for (int j = 0; j < n_kv; ++j) {
        if (!act(it, w, j)) continue;
        const int k_lo = j * TK;
        const int n_col = ((min(TK, L - k_lo) + 15) >> 4) << 4;             // key columns the PV product reads
        // Dropout keep flags of this thread's two 32-column chunks, generated BEFORE waiting for S: the Philox rounds depend
        // only on (b, h, q, key group), so they run in the shadow of the S = Q K^T round trip instead of after it.
        uint32_t kf[2][8];
        if (drop_on && !warp_oob) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int c_lo = k_lo + 32 * ch + 64 * c;
            if (32 * ch + 64 * c < n_col && c_lo < hi_max && c_lo + 31 >= lo_min) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const uint64_t grp = attn_drop_group(it.b, a.nh, it.h, L, qc, (c_lo + 16 * g) >> 4);
                const uint4 rnd = dropout_rand16(a.drop, a.drop_site, grp);
                const uint32_t t4 = dropout_thresh4(a.drop, grp);
                kf[c][4 * g + 0] = keep_flags4(rnd.x, t4);
                kf[c][4 * g + 1] = keep_flags4(rnd.y, t4);
                kf[c][4 * g + 2] = keep_flags4(rnd.z, t4);
                kf[c][4 * g + 3] = keep_flags4(rnd.w, t4);
              }
            }
          }
        }
        mbar_wait(&sh->s_full[w], s_cnt & 1u);
        ++s_cnt;
        tc_fence_after();
        float m_raw = -INFINITY;
        if (!warp_oob) {
          if (!have_prev) {
            // seed the reference exponent from the first 16 keys of the row's first tile (both threads of a row read the
            // same columns, so they agree without an exchange)
            uint32_t v[16];
            tmem_ld16(t_lane + 128 * w, v);
            tmem_ld_wait();
            float mx = -INFINITY, mx_any = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float sv = __uint_as_float(v[i]);
              mx_any = fmaxf(mx_any, sv);
              mx = fmaxf(mx, (static_cast<uint32_t>(k_lo + i - m_lo) < static_cast<uint32_t>(m_hi - m_lo)) ? sv : -INFINITY);
            }
            ref = (mx == -INFINITY ? mx_any : mx) * scale2;
          } else {
            // the row maximum of the PREVIOUS tile over both column halves (written before that tile's p_full arrive, read
            // after this tile's s_full wait: ordered through the two mbarriers): lazy move of the reference exponent
            const float m_prev = fmaxf(sh->xmax[w][(tcount - 1u) & 1u][0][r], sh->xmax[w][(tcount - 1u) & 1u][1][r]);
            if (m_prev > ref + kLazy) {
              const float alpha = ex2(ref - m_prev);
              l_loc *= alpha;
              alpha_pend = alpha;                          // O (all products so far) is rescaled below, before this tile's PV
              ref = m_prev;
            }
          }
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const int cc = 32 * ch + 64 * c;               // first column of the chunk inside the tile: the two threads of a row
                                                           // own INTERLEAVED chunks (ch, ch + 2), so a tile that is half past the
                                                           // sequence end or half masked still splits evenly between them
            if (cc >= n_col) break;
            const int c_lo = k_lo + cc;
            uint32_t pk[16];
            const bool any = c_lo < hi_max && c_lo + 31 >= lo_min;      // hi_max <= L
            if (any) {
              uint32_t v[32];
              tmem_ld32(t_s + 64 * c, v);
              tmem_ld_wait();
              const int rel_lo = m_lo - c_lo, rel_hi = m_hi - c_lo;
              const bool full = warp_whole && lo_max <= c_lo && c_lo + 32 <= hi_min;
              if (full) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
                  m_raw = fmaxf(m_raw, fmaxf(s0, s1));
                  const float p0 = ex2(fmaf(s0, scale2, -ref)), p1 = ex2(fmaf(s1, scale2, -ref));
                  l_loc += p0 + p1;
                  pk[i >> 1] = pack_bf16x2(p0, p1);
                }
              } else {
                const uint32_t span = (rel_hi > rel_lo && q < L) ? static_cast<uint32_t>(rel_hi - rel_lo) : 0u;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float s0 = (static_cast<uint32_t>(i - rel_lo) < span) ? __uint_as_float(v[i]) : -INFINITY;
                  const float s1 = (static_cast<uint32_t>(i + 1 - rel_lo) < span) ? __uint_as_float(v[i + 1]) : -INFINITY;
                  m_raw = fmaxf(m_raw, fmaxf(s0, s1));
                  const float p0 = ex2(fmaf(s0, scale2, -ref)), p1 = ex2(fmaf(s1, scale2, -ref));
                  l_loc += p0 + p1;
                  pk[i >> 1] = pack_bf16x2(p0, p1);
                }
              }
              if (drop_on) {
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                  const uint32_t fl = c == 0 ? kf[0][x] : kf[1][x];
                  pk[2 * x] &= keep_mask_pair(fl, 0);
                  pk[2 * x + 1] &= keep_mask_pair(fl, 1);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = 0u;
            }
            if (c == 0 && have_prev) wait_prev_o();          // the previous PV product has finished reading this P buffer
            store_pk32(sPw, r, ch + 2 * c, pk);
          }
          if (have_prev) {
            if (32 * ch >= n_col) wait_prev_o();             // (this thread stored nothing above)
            if (__any_sync(0xffffffffu, alpha_pend != 1.f)) {
              // rare: some row of this warp moved its reference: rescale its O columns in TMEM (tcgen05.ld / st are
              // warp-collective; rows that did not move multiply by 1).  The previous PV product is complete (waited above).
              uint32_t o[32];
              tmem_ld32(t_o, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_pend);
              tmem_st32(t_o, o);
              tmem_st_wait();
              alpha_pend = 1.f;
            }
          }
          sh->xmax[w][tcount & 1u][ch][r] = m_raw * scale2;
        } else if (have_prev) {
          wait_prev_o();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->p_full[w]);
        ++o_cnt;
        ++tcount;
        have_prev = true;
      }
      if (have_prev) wait_prev_o();
      // row sums: combine the two column halves
      sh->xsum[w][ch][r] = l_loc;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
      if (!warp_oob && have_prev) {
        const float l_row = sh->xsum[w][0][r] + sh->xsum[w][1][r];
        const float inv = l_row > 0.f ? (drop_on ? a.drop.scale : 1.f) / l_row : 0.f;     // deferred dropout keep-scale
        bf16* dst = static_cast<bf16*>(a.ctx) + (static_cast<long>(row0) + qc) * H + it.h * D + 32 * ch;
        uint32_t o[32];
        tmem_ld32(t_o, o);
        tmem_ld_wait();
        if (q < L) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
            reinterpret_cast<uint4*>(dst)[c] = u;
          }
          if (ch == 0) a.lse[(static_cast<long>(it.b) * a.nh + it.h) * L + q] = ref * kLn2 + logf(l_row);
        }
        tc_fence_before();
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");      // xsum is free for the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 18) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

int attention_fwd_legacy_tc05(const AttnArgs& a, cudaStream_t s);

int attention_fwd_tc05(const AttnArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.qkv && a.ctx && a.lse && a.mode && a.t_len, "attention_fwd: null argument");
  static int legacy = -1;                    // MV_ATTN_FWD_LEGACY=1: the r01 kernel (A/B measurements)
  if (legacy < 0) { const char* e = getenv("MV_ATTN_FWD_LEGACY"); legacy = e ? atoi(e) : 0; }
  // byte-sliced threshold compare of the keep bits needs a threshold <= 128 (p <= 0.5)
  if (legacy || (a.drop_on && (a.drop.thresh4 & 0xFFu) > 127u)) return attention_fwd_legacy_tc05(a, s);
  const int H = a.nh * D;
  CUtensorMap tm;
  int rc = tmap_encode_2d(&tm, TMAP_BF16, a.qkv, 3 * H, static_cast<uint64_t>(a.B) * a.L, static_cast<uint64_t>(3 * H) * 2, D, TQ);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    MV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmem));
    attr = true;
  }
  const int n_q = (a.L + TQ - 1) / TQ, n_qp = (n_q + 1) / 2;
  const int n_items = a.B * a.nh * n_qp;
  const int grid = n_items < device_sm_count() ? n_items : device_sm_count();
  MV_CUDA_CHECK(launch_pdl(attn_fwd_ws_kernel, dim3(grid), dim3(kThreads), kWsSmem, s, tm, a, n_items));
  MV_LAUNCH_CHECK();
  return 0;
}

}  // namespace mv
