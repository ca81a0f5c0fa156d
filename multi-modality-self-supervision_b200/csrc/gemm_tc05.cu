// gemm_tc05.cu — persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma (cta_group::1, 128 x BN x 16,
//   fp32 accumulators double-buffered in TMEM) -> tcgen05.ld epilogue (bias / GELU / residual+dropout / tanh /
//   GELU-backward) -> swizzled smem staging -> TMA store (or TMA fp32 reduce-add for split-K weight gradients).
//
// Replaces, for the MedViLL pre-training step, every nn.Linear contraction the reference runs through cuBLAS
// (SURVEY.md §2b K1,K4-K6,K8-K11,K13): Q/K/V, attention-output, FFN, image projection, pooler, MLM transform and
// decoder, in forward / dgrad / wgrad form (see gemm.h for the three operand-major combinations).
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 and 6-9 = two epilogue groups,
// one per TMEM accumulator stage (TMEM lane quarter = warp_id % 4); residual / GELU-backward operand tiles are TMA-prefetched one chunk ahead into the staging ring. Grid = min(#SMs, work units); unit = (m_tile, n_tile, k_split), m fastest so that CTAs
// running concurrently share the weight tile through L2.
#include "gemm.h"
#include "tc05.cuh"
#include "tmap.h"

namespace mv {
using namespace tc05;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 320;   // EW = 4: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 / 6-9 two epilogue groups (EW = 8: 576 threads)
constexpr int kMaxStages = 8;
constexpr uint32_t STG_BYTES = 4096;  // one staging buffer: 32 rows x 128 B (swizzled)
// per epilogue warp: out[2] + second[2], where `second` is either the pre-activation output (EPI_BIAS_GELU) or the
// TMA-loaded epilogue INPUT tile (residual / GELU-backward operand) — the two never coexist
constexpr uint32_t STG_TOTAL = 8 /*epilogue warps*/ * 2 * STG_BYTES;

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_total, kb_per_split, stages;
  int epi;
  const float* bias;
  int has_c2, has_in, accumulate;
  int drop_on; uint32_t drop_site; DropoutCfg drop;
  int a_win, win_wo, win_ho;          // sliding-window A operand (gemm.h): output image width / height
};

struct Barriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t in_bar[8][2];    // per epilogue warp, per input buffer
  uint32_t tmem_base;
};

// read 32 bf16 (columns half*32 .. +32 of row `lane`) from a swizzled [32 x 64] staging tile
__device__ __forceinline__ void unstage_bf16_half(const uint8_t* stg, int lane, int half, float (&out)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(stg + sw128_off(lane, half * 4 + j));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    out[8 * j + 0] = a.x; out[8 * j + 1] = a.y; out[8 * j + 2] = b.x; out[8 * j + 3] = b.y;
    out[8 * j + 4] = c.x; out[8 * j + 5] = c.y; out[8 * j + 6] = d.x; out[8 * j + 7] = d.y;
  }
}

// Epilogue math on 32 consecutive columns [col0, col0+32) of output row `row`.
// f: accumulators in, final values out; pre: pre-activation (EPI_BIAS_GELU); in: residual / GELU-backward operand.
__device__ __forceinline__ void epilogue_apply(const GemmParams& p, float (&f)[32], float (&pre)[32], const float (&in)[32],
                                               int row, int col0, float bias_lane) {
  // bias_lane: bias[col0 + lane] (0 past N), prefetched by the caller one step ahead so that its L2 latency never sits on
  // the accumulator-drain path; column j's value is fetched from lane j.
  const int epi = p.epi;
  if (epi == EPI_BIAS || epi == EPI_BIAS_GELU || epi == EPI_BIAS_RESID || epi == EPI_BIAS_TANH || epi == EPI_BIAS_GELU_GRAD) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += __shfl_sync(0xffffffffu, bias_lane, j);
  }
  if (epi == EPI_BIAS_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) { pre[j] = f[j]; f[j] = gelu_fast(f[j]); }
  } else if (epi == EPI_BIAS_GELU_GRAD) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float cdf, pdf;
      gelu_fast_parts(f[j], cdf, pdf);          // one shared exponential for the activation and its derivative
      pre[j] = fmaf(f[j], pdf, cdf);
      f[j] *= cdf;
    }
  } else if (epi == EPI_BIAS_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = tanhf(f[j]);
  } else if (epi == EPI_BIAS_RESID || epi == EPI_RESID) {
    if (p.drop_on && epi == EPI_BIAS_RESID) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const uint64_t e = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.N) + col0 + 16 * g;
        const uint4 keep = dropout_keep16(p.drop, p.drop_site, e >> 4);
#pragma unroll
        for (int j = 0; j < 16; ++j) f[16 * g + j] = keep16_bit(keep, j) ? f[16 * g + j] * p.drop.scale : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += in[j];
  } else if (epi == EPI_DGELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= gelu_fast_grad(in[j]);
  } else if (epi == EPI_MUL) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= in[j];
  }
}

// write 32 fp32 values as bf16 into half `half` (0/1) of a [32 rows x 64 cols] swizzled staging tile
__device__ __forceinline__ void stage_bf16_half(uint8_t* stg, int lane, int half, const float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
    u.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
    u.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
    u.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
    *reinterpret_cast<uint4*>(stg + sw128_off(lane, half * 4 + j)) = u;
  }
}
// read 32 fp32 (row `lane`) from a swizzled [32 x 32] fp32 staging tile
__device__ __forceinline__ void unstage_f32(const uint8_t* stg, int lane, float (&out)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 u = *reinterpret_cast<const float4*>(stg + sw128_off(lane, j));
    out[4 * j] = u.x; out[4 * j + 1] = u.y; out[4 * j + 2] = u.z; out[4 * j + 3] = u.w;
  }
}
// write 32 fp32 values into a [32 rows x 32 cols] fp32 swizzled staging tile
__device__ __forceinline__ void stage_f32(uint8_t* stg, int lane, const float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<float4*>(stg + sw128_off(lane, j)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  }
}

// ---- 16-column variants for the wide epilogue (EW = 8: two warps per TMEM lane quarter, 113 registers per thread) ----
// bf16 staging tile [32 rows x 64 cols] swizzled; `c16` = which 16-column quarter of the 64 columns
__device__ __forceinline__ void stage_bf16_16(uint8_t* stg, int lane, int c16, const float (&f)[16]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 u;
    u.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
    u.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
    u.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
    u.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
    *reinterpret_cast<uint4*>(stg + sw128_off(lane, c16 * 2 + j)) = u;
  }
}
__device__ __forceinline__ void unstage_bf16_16(const uint8_t* stg, int lane, int c16, float (&out)[16]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(stg + sw128_off(lane, c16 * 2 + j));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    out[8 * j + 0] = a.x; out[8 * j + 1] = a.y; out[8 * j + 2] = b.x; out[8 * j + 3] = b.y;
    out[8 * j + 4] = c.x; out[8 * j + 5] = c.y; out[8 * j + 6] = d.x; out[8 * j + 7] = d.y;
  }
}
// epilogue math on 16 consecutive columns [col0, col0 + 16) of output row `row`; bias of column j comes from lane bias_src + j
__device__ __forceinline__ void epilogue_apply16(const GemmParams& p, float (&f)[16], float (&pre)[16], const float (&in)[16],
                                                 int row, int col0, float bias_lane, int bias_src) {
  const int epi = p.epi;
  if (epi == EPI_BIAS || epi == EPI_BIAS_GELU || epi == EPI_BIAS_RESID || epi == EPI_BIAS_TANH || epi == EPI_BIAS_GELU_GRAD) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] += __shfl_sync(0xffffffffu, bias_lane, bias_src + j);
  }
  if (epi == EPI_BIAS_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { pre[j] = f[j]; f[j] = gelu_fast(f[j]); }
  } else if (epi == EPI_BIAS_GELU_GRAD) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float cdf, pdf;
      gelu_fast_parts(f[j], cdf, pdf);
      pre[j] = fmaf(f[j], pdf, cdf);
      f[j] *= cdf;
    }
  } else if (epi == EPI_BIAS_TANH) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = tanhf(f[j]);
  } else if (epi == EPI_BIAS_RESID || epi == EPI_RESID) {
    if (p.drop_on && epi == EPI_BIAS_RESID) {
      const uint64_t e = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.N) + col0;
      const uint4 keep = dropout_keep16(p.drop, p.drop_site, e >> 4);
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = keep16_bit(keep, j) ? f[j] * p.drop.scale : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] += in[j];
  } else if (epi == EPI_DGELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] *= gelu_fast_grad(in[j]);
  } else if (epi == EPI_MUL) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] *= in[j];
  }
}

// Work unit -> (m tile, n tile, k-block range).  Units are walked by CTA (or CTA pair) c as c, c + stride, ...:
//   plain GEMMs   : n tile fastest, so the CTAs running at the same time share a few row blocks of the (large) activation
//                   operand through L2 while the whole weight matrix (a few MB) stays L2-resident;
//   split-K wgrad : tile fastest / split slowest, so all output tiles walk the same K range together and every byte
//                   of both activation operands is fetched from HBM once.
struct Unit { int m_tile, n_tile, kb0, kb1; };
__device__ __forceinline__ Unit decode_unit(const GemmParams& p, int unit) {
  const int tiles = p.m_tiles * p.n_tiles;
  const int split = unit / tiles;
  const int tile = unit - split * tiles;
  Unit u;
  u.m_tile = tile / p.n_tiles;
  u.n_tile = tile - u.m_tile * p.n_tiles;
  u.kb0 = split * p.kb_per_split;
  u.kb1 = min(p.kb_total, u.kb0 + p.kb_per_split);
  return u;
}

// TWO = true: CTA-pair mode.  Two CTAs of a 2-CTA cluster compute one 256 x BN tile with tcgen05.mma.cta_group::2: CTA r
// loads rows [m0 + 128 r, +128) of A and rows [n0 + r BN/2, +BN/2) of B, owns accumulator rows [128 r, +128) in its own
// TMEM and runs its own epilogue.  Only the leader (rank 0) issues MMAs; every TMA load of the pair is accounted on the
// leader's `full` barrier, MMA completion is multicast to both CTAs' `empty` / `tfull` barriers, and the peer's epilogue
// warps release the accumulator by arriving remotely on the leader's `tempty`.  Each SM then reads only half of B from
// shared memory per MMA — single-CTA tiles are capped at ~60 % of the tensor pipe by shared-memory bandwidth
// (TMA fill + UMMA operand reads, profiles/r01_ncu_gemm_*).
// EW = epilogue warps per accumulator stage.  4: one warp per TMEM lane quarter drains all BN columns (r01).  8: TWO warps per
// lane quarter, each draining one 32-column half of every 64-column chunk into a shared staging tile (pair-synchronised with
// a 64-thread named barrier); used for the transcendental epilogues, where one warp per scheduler could not issue the
// ~26 instructions per element inside two MMA tile-times (profiles/r01_gemm_epilogue_notes.md).
template <int BN, bool A_MN, bool B_MN, bool OUT_F32, bool TWO, int EW>
__global__ void __launch_bounds__(64 + 64 * EW, 1)
gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX,
                 const GemmParams p) {
  // tmX: the second epilogue tensor — pre-activation OUTPUT (has_c2) or residual / aux INPUT (has_in); [M, N] bf16
  constexpr int BNL = TWO ? BN / 2 : BN;          // B rows (or columns) held by THIS CTA
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BNL * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t crank = TWO ? cluster_ctarank() : 0u;
  const int cta_id = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);      // work-unit stream index
  const int cta_stride = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr uint32_t TMEM_COLS = (BN <= 128) ? 256u : 512u;
  constexpr uint32_t ACC_STRIDE = (BN <= 128) ? 128u : 256u;
  constexpr int CHUNK_COLS = OUT_F32 ? 32 : 64;
  constexpr int NCHUNK = BN / CHUNK_COLS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* stg_base = smem + static_cast<uint32_t>(p.stages) * STAGE_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(stg_base + STG_TOTAL);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (p.has_c2 || p.has_in) tma_prefetch_desc(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tfull[a], 1);
      mbar_init(&bars->tempty[a], TWO ? 2 * EW : EW);     // pair mode: local + remote epilogue warps
    }
    for (int w = 0; w < 8; ++w) { mbar_init(&bars->in_bar[w][0], 1); mbar_init(&bars->in_bar[w][1], 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (TWO) { tmem_alloc_pair(&bars->tmem_base, TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(&bars->tmem_base, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_sync();                                       // everything above overlapped the previous kernel's tail

  const int tiles = p.m_tiles * p.n_tiles;          // pair mode: m_tiles counts 256-row tiles
  const int total_units = tiles * p.splits;   // see decode_unit for the order
  constexpr int BMT = TWO ? 2 * BM : BM;            // rows of one work unit

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp walks the loop (warp-uniform control flow keeps addresses / coordinates in uniform registers, which
    // is what UTMALDG consumes); one elected lane arms the barrier and issues the copies.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full0 = TWO ? mapa_cluster(smem_u32(&bars->full[0]), 0) : 0u;   // leader's full[0] (cluster address)
    for (int unit = cta_id; unit < total_units; unit += cta_stride) {
      const Unit u = decode_unit(p, unit);
      const int m0 = u.m_tile * BMT + static_cast<int>(crank) * BM;     // this CTA's 128 rows of A
      const int n0 = u.n_tile * BN + static_cast<int>(crank) * BNL;     // this CTA's share of B
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        mbar_wait(&bars->empty[stage], phase ^ 1u);
        uint8_t* sA = smem + static_cast<uint32_t>(stage) * STAGE_BYTES;
        uint8_t* sB = sA + A_BYTES;
        if (elect_one()) {
          if constexpr (TWO) {
            // the leader arms its barrier for BOTH CTAs' bytes; every load of the pair completes on that barrier
            if (crank == 0) mbar_expect_tx(&bars->full[stage], 2 * STAGE_BYTES);
            const uint32_t fb = full0 + static_cast<uint32_t>(stage) * 8u;
            if constexpr (!A_MN) {
              if (p.a_win) {                        // row m0 = (b, y, x0): window row y + kb of sample b, pixels x0 .. x0 + 127
                const int x0 = m0 % p.win_wo, yb = m0 / p.win_wo;
                tma_load_4d_pair(&tmA, fb, sA, 0, x0, yb % p.win_ho + kb, yb / p.win_ho);
              } else {
                tma_load_2d_pair(&tmA, fb, sA, kb * BK, m0);
              }
            } else {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g) tma_load_2d_pair(&tmA, fb, sA + g * 8192, m0 + g * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d_pair(&tmB, fb, sB, kb * BK, n0);
            } else {
#pragma unroll
              for (int g = 0; g < BNL / 64; ++g) tma_load_2d_pair(&tmB, fb, sB + g * 8192, n0 + g * 64, kb * BK);
            }
          } else {
            mbar_expect_tx(&bars->full[stage], STAGE_BYTES);
            if constexpr (!A_MN) {
              if (p.a_win) {
                const int x0 = m0 % p.win_wo, yb = m0 / p.win_wo;
                tma_load_4d(&tmA, &bars->full[stage], sA, 0, x0, yb % p.win_ho + kb, yb / p.win_ho);
              } else {
                tma_load_2d(&tmA, &bars->full[stage], sA, kb * BK, m0);
              }
            } else {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g) tma_load_2d(&tmA, &bars->full[stage], sA + g * 8192, m0 + g * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d(&tmB, &bars->full[stage], sB, kb * BK, n0);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 64; ++g) tma_load_2d(&tmB, &bars->full[stage], sB + g * 8192, n0 + g * 64, kb * BK);
            }
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop, one elected lane issues tcgen05.mma / tcgen05.commit (the same lane every time: commit tracks
    // the MMAs of the issuing thread).  Pair mode: only the leader CTA issues.
    if (crank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BMT, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint32_t a_addr0 = smem_u32(smem), b_addr0 = a_addr0 + A_BYTES;          // stage 0
      const uint64_t da0 = A_MN ? make_smem_desc_sw128(a_addr0, 8192, 1024) : make_smem_desc_sw128(a_addr0, 16, 1024);
      const uint64_t db0 = B_MN ? make_smem_desc_sw128(b_addr0, 8192, 1024) : make_smem_desc_sw128(b_addr0, 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = cta_id; unit < total_units; unit += cta_stride) {
        const Unit u = decode_unit(p, unit);
        const int kb0 = u.kb0, kb1 = u.kb1;
        mbar_wait(&bars->tempty[acc], acc_phase ^ 1u);      // released by the epilogue (of both CTAs in pair mode)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->full[stage], phase);             // TMA bytes (of both CTAs in pair mode) have landed
          tc_fence_after();
          // Descriptors are built once (da0/db0) and advanced with one 64-bit add per MMA.  The start-address field
          // counts 16-byte units:
          //   K-major : 16 bf16 along K = +32 B inside the 128 B swizzle row          -> +2
          //   MN-major: 16 K rows = +2048 B (8-row K groups 1024 B, 64-wide MN groups 8192 B apart) -> +128
          const uint64_t da = da0 + static_cast<uint64_t>(static_cast<uint32_t>(stage) * (STAGE_BYTES >> 4));
          const uint64_t db = db0 + static_cast<uint64_t>(static_cast<uint32_t>(stage) * (STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t dak = da + static_cast<uint64_t>(k * (A_MN ? 128 : 2)), dbk = db + static_cast<uint64_t>(k * (B_MN ? 128 : 2));
              if constexpr (TWO) umma_bf16_pair(d_tmem, dak, dbk, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, dak, dbk, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs in pair mode) once these MMAs have read it
            if constexpr (TWO) umma_commit_pair(&bars->empty[stage]); else umma_commit(&bars->empty[stage]);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs in pair mode)
        if (elect_one()) {
          if constexpr (TWO) umma_commit_pair(&bars->tfull[acc]); else umma_commit(&bars->tfull[acc]);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if constexpr (EW == 8) {
    // ===================== epilogue warps, two per TMEM lane quarter =====================
    // Group g (accumulator stage g) = warps 2 + 8 g .. 9 + 8 g.  Warps wi and wi + 4 of a group share lane quarter q and one
    // staging slot (2 x 4 KB): warp `hf` drains the 32-column half `hf` of every 64-column chunk, 16 columns at a time.  The
    // half-0 warp ("leader") owns the slot's TMA traffic (input prefetch, stores, bulk-group waits); a 64-thread named barrier
    // orders the two warps around each staging buffer.
    static_assert(!OUT_F32, "the wide epilogue is instantiated for activation-dtype outputs only");
    const int q = warp & 3;
    const int ew = warp - 2;            // 0..15
    const int grp = ew >> 3;            // == accumulator stage
    const int hf = (ew >> 2) & 1;       // column half inside a 64-column chunk
    const int slot = grp * 4 + q;       // staging slot / input barrier pair, 0..7
    const bool leader = hf == 0;
    uint8_t* my_stg = stg_base + static_cast<uint32_t>(slot) * (2 * STG_BYTES);
    uint64_t* in_bar = bars->in_bar[slot];
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory"); };
    const int acc = grp;
    uint32_t acc_phase = 0;
    uint32_t cnt = 0;
    const bool has_in = p.has_in != 0, has_c2 = p.has_c2 != 0;
    auto issue_in = [&](uint32_t chunk_cnt, int m0, int n0, int c) {   // leader lane 0: prefetch the epilogue input tile
      const uint32_t b = chunk_cnt & 1u;
      mbar_expect_tx(&in_bar[b], STG_BYTES);
      tma_load_2d(&tmX, &in_bar[b], my_stg + b * STG_BYTES, n0 + c * 64, m0 + q * 32);
    };
    int it = 0;
    const uint32_t tempty_remote = TWO ? mapa_cluster(smem_u32(&bars->tempty[acc]), 0) : 0u;
    for (int unit = cta_id; unit < total_units; unit += cta_stride, ++it) {
      if ((it & 1) != grp) continue;
      const Unit u = decode_unit(p, unit);
      const int m0 = u.m_tile * BMT + static_cast<int>(crank) * BM;
      const int n0 = u.n_tile * BN;
      if (has_in && leader && lane == 0) {
        tma_wait_group_read<0>();
        issue_in(cnt, m0, n0, 0);
      }
      const bool has_bias = p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU || p.epi == EPI_BIAS_RESID || p.epi == EPI_BIAS_TANH || p.epi == EPI_BIAS_GELU_GRAD;
      auto bias_at = [&](int col) { return (has_bias && col < p.N) ? __ldg(p.bias + col) : 0.f; };
      float bias_next = bias_at(n0 + hf * 32 + lane);       // this warp's 32 columns of chunk 0
      mbar_wait(&bars->tfull[acc], acc_phase);
      acc_phase ^= 1u;
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_row = tmem_base + static_cast<uint32_t>(acc) * ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
      const int nch = min(BN / 64, (p.N - n0 + 63) / 64);     // chunks that hold real columns (N may end inside the tile)
#pragma unroll 1
      for (int c = 0; c < nch; ++c, ++cnt) {
        const uint32_t buf = has_c2 ? 0u : (cnt & 1u);
        if (leader && lane == 0) {
          if (has_in) {
            if (c + 1 < nch) {
              tma_wait_group_read<0>();
              issue_in(cnt + 1, m0, n0, c + 1);
            }
          } else if (has_c2) {
            tma_wait_group_read<0>();
          } else {
            tma_wait_group_read<1>();
          }
        }
        pair_sync();                                   // staging buffer `buf` is free for both warps
        uint8_t* s1 = my_stg + buf * STG_BYTES;
        uint8_t* s2 = has_c2 ? my_stg + STG_BYTES : s1;
        if (has_in) mbar_wait(&in_bar[buf], (cnt >> 1) & 1u);
        const float bias_cur = bias_next;
        bias_next = bias_at(n0 + (c + 1) * 64 + hf * 32 + lane);
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t v[16];
          tmem_ld16(t_row + static_cast<uint32_t>(c * 64 + hf * 32 + sub * 16), v);
          tmem_ld_wait();
          float f[16], pre[16], in[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (has_in) unstage_bf16_16(s2, lane, hf * 2 + sub, in);
          epilogue_apply16(p, f, pre, in, row, n0 + c * 64 + hf * 32 + sub * 16, bias_cur, sub * 16);
          stage_bf16_16(s1, lane, hf * 2 + sub, f);
          if (has_c2) stage_bf16_16(s2, lane, hf * 2 + sub, pre);
        }
        if (c == nch - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (TWO) mbar_arrive_cluster(tempty_remote);
            else mbar_arrive(&bars->tempty[acc]);
          }
        }
        fence_proxy_async_smem();
        pair_sync();                                   // both halves of the chunk are staged
        if (leader && lane == 0) {
          const int r0 = m0 + q * 32;
          const int c0 = n0 + c * 64;
          if (r0 < p.M && c0 < p.N) {
            tma_store_2d(&tmC, s1, c0, r0);
            if (has_c2) tma_store_2d(&tmX, s2, c0, r0);
          }
          tma_commit_group();
        }
      }
    }
    if (leader && lane == 0) tma_wait_group<0>();
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // Two groups of four warps (2-5 and 6-9).  Group g drains accumulator stage g, i.e. every other work unit of this
    // CTA, so two tiles' epilogues overlap each other and the MMA warp's next mainloop.  Staging per warp: 2 x 4 KB.
    //   plain epilogues : out double-buffered
    //   residual / dGELU: the TMA-prefetched INPUT tile is overwritten in place by the output, double-buffered
    //   GELU + pre-act  : buffer 0 = output, buffer 1 = pre-activation (single-buffered; the other group fills the gap)
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    const int ew = warp - 2;            // epilogue warp index 0..7
    const int grp = ew >> 2;            // == accumulator stage
    uint8_t* my_stg = stg_base + static_cast<uint32_t>(ew) * (2 * STG_BYTES);
    uint64_t* in_bar = bars->in_bar[ew];
    const int acc = grp;
    uint32_t acc_phase = 0;
    uint32_t cnt = 0;            // chunks processed so far by this warp: buffer = cnt & 1, input parity = (cnt >> 1) & 1
    const bool has_in = p.has_in != 0, has_c2 = p.has_c2 != 0;
    auto issue_in = [&](uint32_t chunk_cnt, int m0, int n0, int c) {   // lane 0 only: prefetch the epilogue input tile
      const uint32_t b = chunk_cnt & 1u;
      mbar_expect_tx(&in_bar[b], STG_BYTES);
      tma_load_2d(&tmX, &in_bar[b], my_stg + b * STG_BYTES, n0 + c * CHUNK_COLS, m0 + q * 32);
    };
    int it = 0;
    const uint32_t tempty_remote = TWO ? mapa_cluster(smem_u32(&bars->tempty[acc]), 0) : 0u;   // leader's barrier
    for (int unit = cta_id; unit < total_units; unit += cta_stride, ++it) {
      if ((it & 1) != grp) continue;
      const Unit u = decode_unit(p, unit);
      const int m0 = u.m_tile * BMT + static_cast<int>(crank) * BM;      // this CTA's 128 accumulator rows
      const int n0 = u.n_tile * BN;                                      // ... over all BN columns
      if (has_in && lane == 0) {       // overlaps the wait for the accumulator
        tma_wait_group_read<0>();      // every store that read my two buffers has drained
        issue_in(cnt, m0, n0, 0);
      }
      const bool has_bias = p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU || p.epi == EPI_BIAS_RESID || p.epi == EPI_BIAS_TANH || p.epi == EPI_BIAS_GELU_GRAD;
      auto bias_at = [&](int col) { return (has_bias && col < p.N) ? __ldg(p.bias + col) : 0.f; };
      float bias_next = bias_at(n0 + lane);      // first 32 columns; later ones are fetched one step ahead
      mbar_wait(&bars->tfull[acc], acc_phase);
      acc_phase ^= 1u;
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_row = tmem_base + static_cast<uint32_t>(acc) * ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
      // chunks that hold real columns: N may end inside the tile (N = 64 of the stem convolution under BN = 128 drains half)
      const int nch = min(NCHUNK, (p.N - n0 + CHUNK_COLS - 1) / CHUNK_COLS);
#pragma unroll 1
      for (int c = 0; c < nch; ++c, ++cnt) {
        const uint32_t buf = has_c2 ? 0u : (cnt & 1u);
        if (lane == 0) {
          if (has_in) {
            if (c + 1 < nch) {
              tma_wait_group_read<0>();              // the other buffer's last store has been read out
              issue_in(cnt + 1, m0, n0, c + 1);      // input tile of the next chunk
            }
          } else if (has_c2) {
            tma_wait_group_read<0>();
          } else {
            tma_wait_group_read<1>();                // out[buf] (stored two chunks ago) is free again
          }
        }
        __syncwarp();
        uint8_t* s1 = my_stg + buf * STG_BYTES;
        uint8_t* s2 = has_c2 ? my_stg + STG_BYTES : s1;     // pre-activation out, or the in-place input tile
        if (has_in) mbar_wait(&in_bar[buf], (cnt >> 1) & 1u);
        if constexpr (!OUT_F32) {
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld32(t_row + static_cast<uint32_t>(c * 64 + half * 32), v);
            tmem_ld_wait();
            float f[32], pre[32], in[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (has_in) unstage_bf16_half(s2, lane, half, in);
            const float bias_cur = bias_next;
            bias_next = bias_at(n0 + c * 64 + half * 32 + 32 + lane);
            epilogue_apply(p, f, pre, in, row, n0 + c * 64 + half * 32, bias_cur);
            stage_bf16_half(s1, lane, half, f);
            if (p.has_c2) stage_bf16_half(s2, lane, half, pre);
          }
        } else {
          uint32_t v[32];
          tmem_ld32(t_row + static_cast<uint32_t>(c * 32), v);
          tmem_ld_wait();
          float f[32], pre[32], in[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (has_in) unstage_f32(s2, lane, in);       // fp32 residual stream: the input tile is fp32 too, replaced in place
          const float bias_cur = bias_next;
          bias_next = bias_at(n0 + c * 32 + 32 + lane);
          epilogue_apply(p, f, pre, in, row, n0 + c * 32, bias_cur);
          stage_f32(s1, lane, f);
        }
        if (c == nch - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (TWO) mbar_arrive_cluster(tempty_remote);     // both CTAs' warps release the leader's barrier
            else mbar_arrive(&bars->tempty[acc]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int r0 = m0 + q * 32;
          const int c0 = n0 + c * CHUNK_COLS;
          if (r0 < p.M && c0 < p.N) {
            if (p.accumulate) tma_reduce_add_2d(&tmC, s1, c0, r0);
            else tma_store_2d(&tmC, s1, c0, r0);
            if (p.has_c2) tma_store_2d(&tmX, s2, c0, r0);
          }
          tma_commit_group();
        }
      }
    }
    if (lane == 0) tma_wait_group<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();   // the peer's smem / barriers / TMEM stay alive until the leader's last MMA retired
  if (warp == 1) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, bool TWO>
constexpr uint32_t stage_bytes() { return BM * BK * 2 + (TWO ? BN / 2 : BN) * BK * 2; }

inline int pick_stages(uint32_t stage_b) {
  const uint32_t budget = 232448u - 1024u - STG_TOTAL - static_cast<uint32_t>(sizeof(Barriers)) - 64u;
  int s = static_cast<int>(budget / stage_b);
  if (s > kMaxStages) s = kMaxStages;
  static int cap = -1;                       // MV_GEMM_STAGES=n caps the ring depth (pipeline-depth experiments)
  if (cap < 0) { const char* e = getenv("MV_GEMM_STAGES"); cap = e ? atoi(e) : 0; }
  if (cap > 0 && s > cap) s = cap;
  return s;
}

template <int BN, bool A_MN, bool B_MN, bool OUT_F32, bool TWO, int EW = 4>
int launch(const GemmDesc& d, GemmParams p, cudaStream_t stream) {
  constexpr int kThreadsEw = 64 + 64 * EW;
  constexpr int BNL = TWO ? BN / 2 : BN;     // B rows / columns one CTA loads per k-block
  constexpr int BMT = TWO ? 2 * BM : BM;     // rows of one work unit
  CUtensorMap tmA, tmB, tmC, tmX;
  int rc;
  if (d.a_win) {
    const uint64_t dims[4] = {64, static_cast<uint64_t>(d.win_wo), static_cast<uint64_t>(d.win_hs), static_cast<uint64_t>(d.win_b)};
    const uint64_t str[3] = {static_cast<uint64_t>(d.win_c) * 2, static_cast<uint64_t>(d.win_ws) * d.win_c * 2,
                             static_cast<uint64_t>(d.win_hs) * d.win_ws * d.win_c * 2};
    rc = tmap_encode_4d(&tmA, TMAP_BF16, d.A, dims, str, BK, BM);
  } else if (!d.a_mn) rc = tmap_encode_2d(&tmA, TMAP_BF16, d.A, d.K, d.M, d.lda * 2, BK, BM);
  else rc = tmap_encode_2d(&tmA, TMAP_BF16, d.A, d.M, d.K, d.lda * 2, 64, BK);
  if (rc) return rc;
  if (!d.b_mn) rc = tmap_encode_2d(&tmB, TMAP_BF16, d.B, d.K, d.N, d.ldb * 2, BK, BNL);
  else rc = tmap_encode_2d(&tmB, TMAP_BF16, d.B, d.N, d.K, d.ldb * 2, 64, BK);
  if (rc) return rc;
  if (OUT_F32) rc = tmap_encode_2d(&tmC, TMAP_F32, d.C, d.N, d.M, d.ldc * 4, 32, 32);
  else rc = tmap_encode_2d(&tmC, TMAP_BF16, d.C, d.N, d.M, d.ldc * 2, 64, 32);
  if (rc) return rc;
  tmX = tmC;
  if (d.C2) {
    rc = tmap_encode_2d(&tmX, TMAP_BF16, d.C2, d.N, d.M, d.ldc2 * 2, 64, 32);
  } else if ((d.epi == EPI_BIAS_RESID || d.epi == EPI_RESID) && d.resid_f32) {
    rc = tmap_encode_2d(&tmX, TMAP_F32, d.resid, d.N, d.M, d.ldr * 4, 32, 32);
  } else if (d.epi == EPI_BIAS_RESID || d.epi == EPI_RESID) {
    rc = tmap_encode_2d(&tmX, TMAP_BF16, d.resid, d.N, d.M, d.ldr * 2, 64, 32);
  } else if (d.epi == EPI_DGELU || d.epi == EPI_MUL) {
    rc = tmap_encode_2d(&tmX, TMAP_BF16, d.aux, d.N, d.M, d.ldaux * 2, 64, 32);
  }
  if (rc) return rc;
  p.m_tiles = (d.M + BMT - 1) / BMT;
  p.n_tiles = (d.N + BN - 1) / BN;
  p.stages = pick_stages(stage_bytes<BN, TWO>());
  // split-K only for fp32 reduce-add outputs (weight gradients): fill ~2 waves of SMs (or SM pairs)
  const int slots = TWO ? gemm_sm_budget() / 2 : gemm_sm_budget();
  const int tiles = p.m_tiles * p.n_tiles;
  int splits = 1;
  if (d.accumulate) {
    // split-K (fp32 reduce-add outputs): pick the split count that minimises waves x (k-blocks per unit + fixed cost)
    if (d.splitk > 0) {
      splits = d.splitk;
    } else {
      const int max_splits = (p.kb_total + 7) / 8;      // keep >= 8 k-blocks per unit
      long best = -1;
      for (int sp = 1; sp <= max_splits && sp <= 64; ++sp) {
        const int kbs = (p.kb_total + sp - 1) / sp;
        const int real = (p.kb_total + kbs - 1) / kbs;
        const long waves = (static_cast<long>(tiles) * real + slots - 1) / slots;
        const long cost = waves * (kbs + 8);            // ~8 k-blocks worth of prologue + reduce-add epilogue per unit
        if (best < 0 || cost < best) { best = cost; splits = sp; }
      }
    }
    if (splits < 1) splits = 1;
    if (splits > p.kb_total) splits = p.kb_total;
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  const int units = tiles * p.splits;
  const int grid = (units < slots ? units : slots) * (TWO ? 2 : 1);
  const uint32_t smem = 1024u + static_cast<uint32_t>(p.stages) * stage_bytes<BN, TWO>() + STG_TOTAL + sizeof(Barriers) + 64u;
  auto kern = gemm_tc05_kernel<BN, A_MN, B_MN, OUT_F32, TWO, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    MV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  if constexpr (TWO) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(kThreadsEw, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmX, p));
  } else {
    MV_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kThreadsEw), smem, stream, tmA, tmB, tmC, tmX, p));
  }
  MV_LAUNCH_CHECK();
  return 0;
}

// MV_GEMM_EW=4 / 8 forces the epilogue width (A/B measurements); default: 8 for the transcendental epilogues
int ew_env() {
  static int v = -2;
  if (v == -2) { const char* e = getenv("MV_GEMM_EW"); v = e ? atoi(e) : 0; }
  return v;
}

template <int BN, bool TWO>
int dispatch_major(const GemmDesc& d, const GemmParams& p, cudaStream_t s) {
  bool wide = d.epi == EPI_BIAS_GELU || d.epi == EPI_BIAS_GELU_GRAD || d.epi == EPI_DGELU;
  if (ew_env() == 4) wide = false; else if (ew_env() == 8) wide = true;
  wide = wide && !d.c_f32 && !d.a_mn;
  if (!d.a_mn && !d.b_mn) {
    if (wide) return launch<BN, false, false, false, TWO, 8>(d, p, s);
    return d.c_f32 ? launch<BN, false, false, true, TWO>(d, p, s) : launch<BN, false, false, false, TWO>(d, p, s);
  }
  if constexpr (!TWO || (BN / 2) % 64 == 0) {      // MN-major B is loaded in 64-column groups
    if (!d.a_mn && d.b_mn && wide) return launch<BN, false, true, false, TWO, 8>(d, p, s);
    if (!d.a_mn && d.b_mn) return d.c_f32 ? launch<BN, false, true, true, TWO>(d, p, s) : launch<BN, false, true, false, TWO>(d, p, s);
    if (d.a_mn && d.b_mn) return d.c_f32 ? launch<BN, true, true, true, TWO>(d, p, s) : launch<BN, true, true, false, TWO>(d, p, s);
  }
  set_error("gemm_bf16_tc05: operand-major combination a_mn=%d,b_mn=%d is not instantiated for this tile", d.a_mn, d.b_mn);
  return -1;
}

// pick the N tile that minimises (waves x per-tile cost) on `slots` SMs (or SM pairs)
int pick_bn(int m_tiles, int N, int slots, bool accumulate, bool pair, bool b_mn) {
  const int cands[3] = {192, 256, 128};
  int best = 0;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (pair && b_mn && (bn / 2) % 64 != 0) continue;
    const long tiles = static_cast<long>(m_tiles) * ((N + bn - 1) / bn);
    const long waves = accumulate ? 1 : (tiles + slots - 1) / slots;
    const double per_tile = bn + 24.0;  // MMA time ~ BN, plus fixed prologue/epilogue overhead
    const double wasted = static_cast<double>(((N + bn - 1) / bn) * bn) / N;
    const double cost = accumulate ? wasted * per_tile / bn : waves * per_tile;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// MV_GEMM_PAIR=0/1 overrides the automatic choice (A/B measurements)
int pair_env() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("MV_GEMM_PAIR");
    v = e ? atoi(e) : -1;
  }
  return v;
}

}  // namespace

int gemm_bf16_tc05(const GemmDesc& d, cudaStream_t stream) {
  MV_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "gemm: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
  MV_REQUIRE(d.A && d.B && d.C, "gemm: null operand");
  MV_REQUIRE(!d.accumulate || d.c_f32, "gemm: accumulate requires fp32 output");
  MV_REQUIRE(!(d.C2 && d.c_f32), "gemm: pre-activation output only with activation-dtype C");
  MV_REQUIRE(!d.C2 || ((d.epi == EPI_BIAS_GELU || d.epi == EPI_BIAS_GELU_GRAD) && d.N % 32 == 0 && d.ldc2 % 8 == 0),
             "gemm: C2 needs EPI_BIAS_GELU / EPI_BIAS_GELU_GRAD, N %% 32 == 0");
  MV_REQUIRE(d.epi != EPI_BIAS_GELU_GRAD || d.C2, "gemm: EPI_BIAS_GELU_GRAD needs C2 (the derivative output)");
  if (d.epi == EPI_BIAS || d.epi == EPI_BIAS_GELU || d.epi == EPI_BIAS_RESID || d.epi == EPI_BIAS_TANH || d.epi == EPI_BIAS_GELU_GRAD)
    MV_REQUIRE(d.bias != nullptr, "gemm: epilogue %d needs bias", d.epi);
  if (d.epi == EPI_BIAS_RESID || d.epi == EPI_RESID)
    MV_REQUIRE(d.resid != nullptr && d.N % 32 == 0 && d.ldr % 8 == 0 && (d.c_f32 != 0) == (d.resid_f32 != 0) && !(d.c_f32 && d.accumulate),
               "gemm: residual epilogue needs resid, N%%32==0, and resid / C of the same type (bf16, or fp32 with resid_f32)");
  if (d.epi == EPI_DGELU || d.epi == EPI_MUL)
    MV_REQUIRE(d.aux != nullptr && d.N % 32 == 0 && d.ldaux % 8 == 0 && !d.c_f32, "gemm: DGELU / MUL epilogue needs aux, N%%32==0, bf16 out");
  GemmParams p;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.m_tiles = (d.M + BM - 1) / BM;
  p.n_tiles = 0;
  p.kb_total = (d.K + BK - 1) / BK;
  p.splits = 1; p.kb_per_split = p.kb_total; p.stages = 0;
  p.epi = d.epi;
  p.bias = d.bias;
  p.has_c2 = d.C2 != nullptr; p.accumulate = d.accumulate;
  p.has_in = (d.epi == EPI_BIAS_RESID || d.epi == EPI_RESID || d.epi == EPI_DGELU || d.epi == EPI_MUL) ? 1 : 0;
  p.drop_on = d.drop_on; p.drop_site = d.drop_site; p.drop = d.drop;
  p.a_win = d.a_win; p.win_wo = d.win_wo; p.win_ho = d.win_ho;
  if (d.a_win) {
    MV_REQUIRE(!d.a_mn && !d.b_mn && !d.accumulate && d.win_c > 0 && 64 % d.win_c == 0 && (d.win_c * 2) % 16 == 0 && d.K % 64 == 0 &&
                   d.win_wo % BM == 0 && d.win_ho + d.K / 64 - 1 <= d.win_hs && d.win_wo + 64 / d.win_c - 1 <= d.win_ws &&
                   static_cast<long>(d.win_b) * d.win_ho * d.win_wo == d.M,
               "gemm: sliding-window A needs a_mn = b_mn = 0, K %% 64 == 0, win_wo %% 128 == 0 and a window inside the input");
  }
  bool pair = d.M > BM;                       // a single 128-row tile gains nothing from a partner SM
  if (pair_env() >= 0) pair = pair_env() != 0;
  if (d.pair >= 0) pair = d.pair != 0;
  const int sms = gemm_sm_budget();
  const int m_tiles = pair ? (d.M + 2 * BM - 1) / (2 * BM) : p.m_tiles;
  int bn = d.bn;
  if (bn == 0) bn = pick_bn(m_tiles, d.N, pair ? sms / 2 : sms, d.accumulate != 0, pair, d.b_mn != 0);
  MV_REQUIRE(bn == 128 || bn == 192 || bn == 256, "gemm: N tile %d not instantiated", bn);
  if (pair) {
    switch (bn) {
      case 128: return dispatch_major<128, true>(d, p, stream);
      case 256: return dispatch_major<256, true>(d, p, stream);
      default: return dispatch_major<192, true>(d, p, stream);
    }
  }
  switch (bn) {
    case 128: return dispatch_major<128, false>(d, p, stream);
    case 256: return dispatch_major<256, false>(d, p, stream);
    default: return dispatch_major<192, false>(d, p, stream);
  }
}

}  // namespace mv
