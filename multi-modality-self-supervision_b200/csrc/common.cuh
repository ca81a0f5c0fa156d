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

// ---- dropout RNG: Philox4x32-10, one call -> 8 x 16-bit lanes ----
// Keyed by (seed, site); counter = 64-bit element-group index (8 consecutive elements of a row).
// The same (seed, site, group) regenerates the same keep-mask in forward and backward, so no mask
// tensor is ever stored.
struct DropoutCfg {
  float p;             // drop probability as configured
  uint32_t thresh16;   // drop iff r16 < thresh16
  float scale;         // 1 / (1 - thresh16/65536)
  uint64_t seed;       // per-step seed
};

__host__ __device__ inline DropoutCfg make_dropout(float p, uint64_t seed) {
  DropoutCfg d;
  d.p = p;
  uint32_t t = static_cast<uint32_t>(p * 65536.0f + 0.5f);
  if (t > 65535u) t = 65535u;
  d.thresh16 = t;
  d.scale = 1.0f / (1.0f - static_cast<float>(t) / 65536.0f);
  d.seed = seed;
  return d;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep-mask (bit i = element i kept) for the 8 consecutive elements of group `g` at dropout site `site`
__device__ __forceinline__ uint32_t dropout_keep8(const DropoutCfg& d, uint32_t site, uint64_t g) {
  const uint4 r = philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), site, 0x4d56u,
                                static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= d.thresh16) << 0; m |= ((r.x >> 16) >= d.thresh16) << 1;
  m |= ((r.y & 0xFFFFu) >= d.thresh16) << 2; m |= ((r.y >> 16) >= d.thresh16) << 3;
  m |= ((r.z & 0xFFFFu) >= d.thresh16) << 4; m |= ((r.z >> 16) >= d.thresh16) << 5;
  m |= ((r.w & 0xFFFFu) >= d.thresh16) << 6; m |= ((r.w >> 16) >= d.thresh16) << 7;
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
