// common.cuh — shared device helpers: error plumbing, counter-based dropout RNG, GELU, bf16 packing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

typedef __nv_bfloat16 bf16;

namespace mv {

// ---- error plumbing: every C-ABI entry returns 0 or a negative status; message via mv_last_error ----
void set_error(const char* fmt, ...);
const char* last_error();
int device_sm_count();
void set_reserved_sms(int n); // persistent GEMM grids use device_sm_count() - n SMs (n SMs stay free for NCCL CTAs)
int gemm_sm_budget();
void count_launch();          // every kernel launch of this library bumps one process-wide counter
long launch_count();

#define MV_CUDA_CHECK(expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      mv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);    \
      return -2;                                                                                    \
    }                                                                                               \
  } while (0)

// after every <<<...>>>: count the launch, surface launch-configuration errors
#define MV_LAUNCH_CHECK()                        \
  do {                                           \
    mv::count_launch();                          \
    MV_CUDA_CHECK(cudaGetLastError());           \
  } while (0)

// ---- programmatic dependent launch (PDL) ----
// Kernels that begin with pdl_wait() may be launched with launch_pdl(): the grid is allowed to start (CTA launch, barrier /
// TMEM / descriptor set-up) while the previous kernel of the stream is still draining, and blocks in pdl_wait() until that
// kernel has completed and its writes are visible.  pdl_trigger() lets the NEXT kernel do the same with respect to this one.
// Rule: a kernel launched through launch_pdl() executes pdl_wait() in every CTA before its first global-memory access (reads
// AND writes: workspaces are reused).  Without the launch attribute both instructions are no-ops.  MV_PDL=0 disables it.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define MV_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      mv::set_error(__VA_ARGS__);      \
      return -1;                       \
    }                                  \
  } while (0)

// ---- math ----
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// d/dx [x * Phi(x)] = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Fast exact-form GELU for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below
// bf16 resolution), sharing ONE exponential between erf(x/sqrt2) and the Gaussian pdf: exp(-(x/sqrt2)^2) == exp(-x^2/2).
// The fp32 check-mode kernels keep erff().
__device__ __forceinline__ void gelu_fast_parts(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));   // exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);            // erf(|x|/sqrt2)
  cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_fast(float x) { float c, p; gelu_fast_parts(x, c, p); return x * c; }
__device__ __forceinline__ float gelu_fast_grad(float x) { float c, p; gelu_fast_parts(x, c, p); return fmaf(x, p, c); }

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- dropout RNG: Philox4x32-7, one call -> 16 x 8-bit lanes ----
// Keyed by (seed, site); counter = 64-bit element-group index (16 consecutive elements of a row).
// The same (seed, site, group) regenerates the same keep-mask in forward and backward, so no mask tensor is ever
// stored.  An element is dropped iff its random byte < the group's threshold.  256 p is not an integer in general
// (p = 0.1 -> 25.6), so the threshold is DITHERED per group: floor(256 p) + 1 with probability frac(256 p), else
// floor(256 p), decided by an 8-bit hash of the group index and the seed (dropout_thresh4).  Every element is then
// dropped with probability p to within 2^-16 (p = 0.1 -> 0.100006; plain 8-bit rounding gave 0.1016) and the keep-scale
// is the reference's 1 / (1 - p).
struct DropoutCfg {
  float p;             // drop probability as configured
  uint32_t thresh4;    // floor(256 p) replicated in 4 bytes
  uint32_t frac8;      // round(256 * frac(256 p)): groups whose hash byte is below it use threshold floor(256 p) + 1
  float scale;         // 1 / (1 - p)
  uint64_t seed;       // per-step seed
};

__host__ __device__ inline DropoutCfg make_dropout(float p, uint64_t seed) {
  DropoutCfg d;
  d.p = p;
  const float t256 = p * 256.0f;
  uint32_t t = static_cast<uint32_t>(t256);
  if (t > 254u) t = 254u;
  uint32_t f = static_cast<uint32_t>((t256 - static_cast<float>(t)) * 256.0f + 0.5f);
  if (f > 255u) { f = 0u; t += 1u; }
  d.thresh4 = t * 0x01010101u;
  d.frac8 = f;
  d.scale = 1.0f / (1.0f - p);
  d.seed = seed;
  return d;
}
// the (replicated) byte threshold of element group g
__device__ __forceinline__ uint32_t dropout_thresh4(const DropoutCfg& d, uint64_t g) {
  const uint32_t h = (static_cast<uint32_t>(g) * 0x9E3779B1u) ^ (static_cast<uint32_t>(g >> 32) * 0x85EBCA6Bu) ^ static_cast<uint32_t>(d.seed >> 7);
  return ((h * 0xC2B2AE35u) >> 24) < d.frac8 ? d.thresh4 + 0x01010101u : d.thresh4;
}

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0, n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1, n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep-mask for the 16 consecutive elements of group `g` at dropout site `site`: byte i of the result (word i/4,
// byte i%4) is 0xFF iff element i is kept.
__device__ __forceinline__ uint4 dropout_keep16(const DropoutCfg& d, uint32_t site, uint64_t g) {
  uint4 r = philox4x32_7(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), site, 0x4d56u,
                         static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
  const uint32_t t4 = dropout_thresh4(d, g);
  r.x = __vcmpgeu4(r.x, t4); r.y = __vcmpgeu4(r.y, t4);
  r.z = __vcmpgeu4(r.z, t4); r.w = __vcmpgeu4(r.w, t4);
  return r;
}
// The raw 128 random bits of group `g` (16 x 8-bit lanes): dropout_keep16 == byte-wise (lane >= threshold) of this.
__device__ __forceinline__ uint4 dropout_rand16(const DropoutCfg& d, uint32_t site, uint64_t g) {
  return philox4x32_7(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), site, 0x4d56u, static_cast<uint32_t>(d.seed),
                      static_cast<uint32_t>(d.seed >> 32));
}
// Keep FLAGS of four 8-bit lanes: bit 7 of byte i is set iff lane i >= threshold (the same decision as __vcmpgeu4, whose
// software emulation costs ~10 instructions per word; this is 3).  Valid for thresholds <= 128, i.e. p <= 0.5:
//   lane >= 128: bit 7 of r itself;  lane < 128: (lane | 0x80) - T = 128 + lane - T has bit 7 set iff lane >= T, and never
//   borrows from the next byte because 128 + lane - T >= 128 - T >= 0.
__device__ __forceinline__ uint32_t keep_flags4(uint32_t r, uint32_t thresh4) { return ((r | 0x80808080u) - thresh4) | r; }
// any threshold: thresholds above 128 (p > 0.5) take the emulated byte compare, whose 0xFF / 0x00 lanes are valid flags too
__device__ __forceinline__ uint32_t keep_flags4_any(uint32_t r, uint32_t thresh4) {
  return (thresh4 & 0xFFu) <= 128u ? keep_flags4(r, thresh4) : __vcmpgeu4(r, thresh4);
}
// Expand the flags of lanes (2j, 2j+1) of `flags` into a bf16x2 AND-mask: PRMT with sign replication (selector bit 3)
// copies bit 7 of the selected byte into all eight bits of the destination byte.
__device__ __forceinline__ uint32_t keep_mask_pair(uint32_t flags, int j) {     // j = 0: lanes 0,1;  j = 1: lanes 2,3
  uint32_t m;                          // inline PTX: the __byte_perm intrinsic documents only selector bits [2:0] of each nibble
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(flags), "r"(0u), "r"(j ? 0xBBAAu : 0x9988u));
  return m;
}

__device__ __forceinline__ bool keep16_bit(const uint4& m, int i) {   // i in [0,16)
  const uint32_t w = i < 4 ? m.x : (i < 8 ? m.y : (i < 12 ? m.z : m.w));
  return (w >> (8 * (i & 3))) & 1u;
}
// keep bits (bit j = element j kept) of the 8-element half `half` (0/1) of a 16-group
__device__ __forceinline__ uint32_t keep16_half_bits(const uint4& m, int half) {
  const uint32_t a = half ? m.z : m.x, b = half ? m.w : m.y;
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { bits |= ((a >> (8 * j)) & 1u) << j; bits |= ((b >> (8 * j)) & 1u) << (4 + j); }
  return bits;
}
// 8 consecutive elements starting at flat element index e (e % 8 == 0): bit j = element j kept
__device__ __forceinline__ uint32_t dropout_keep8(const DropoutCfg& d, uint32_t site, uint64_t e) {
  return keep16_half_bits(dropout_keep16(d, site, e >> 4), static_cast<int>((e >> 3) & 1));
}

// all-ones / zero 32-bit mask of lane j (0..3) of `flags` (for fp32 operands)
__device__ __forceinline__ uint32_t keep_mask_elem(uint32_t flags, int j) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(flags), "r"(0u), "r"(0x8888u + 0x1111u * static_cast<uint32_t>(j)));
  return m;
}

// ---- warp / block reductions ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mv
