// bn.cu — BatchNorm2d for the frozen ResNet-50 trunk on channels-last activations [rows = B*H*W, C] (bf16 or fp32):
// train-mode batch statistics (models/train_origin.py:72 runs model.train() on frozen weights) + affine + optional
// residual add + ReLU in one apply pass.  The convolutions stay on cuDNN (BASELINE.json north_star (4)); these two
// HBM-bound passes replace PyTorch's native channels-last batch-norm kernels, which ran at ~1/7 of the HBM roofline
// (profiles/r01_launches_step.txt).  Algorithmic traffic: stats 1 read, apply 1 read (+1 residual read) + 1 write.
//
// Numerics: per-thread fp32 sum / sum-of-squares over <= a few thousand rows, converted to (n, mean, M2) and merged with
// Chan's parallel update across threads and CTAs; running stats use the unbiased variance (PyTorch semantics).
#include "kernels.h"

namespace mv {
namespace {

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void st8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
  if (nb == 0.f) return;
  const float nt = n + nb, d = meanb - mean;
  mean += d * (nb / nt);
  m2 += m2b + d * d * (n * nb / nt);
  n = nt;
}

// Statistics pass.  grid = (row parts, 64-channel blocks).  256 threads = 32 row lanes x 8 chunk threads: a warp reads four
// 128-byte row segments per iteration (full sectors), a CTA walks its row slab 32 rows at a time.  Every thread keeps
// fp32 sum / sum-of-squares of 8 channels, converts to (n, mean, M2) and the 32 row lanes are merged with Chan updates
// through shared memory.  Partials: [channel block][part][64 channels][3] — per layer at most 148*4*64*3 floats, so the
// finalize pass reads a few hundred KB instead of the nparts*C*3 table the first version wrote (14.5 MB at C = 2048).
constexpr int kBnCh = 64;        // channels per CTA
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, long rows, int C, long rows_per_cta,
                                                       float* __restrict__ partial) {
  __shared__ float red[32][kBnCh][3];      // 24 KB
  pdl_sync();
  const int ct = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.y * kBnCh + ct * 8;
  const long r0 = static_cast<long>(blockIdx.x) * rows_per_cta;
  const long r1 = min(rows, r0 + rows_per_cta);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float cnt = 0.f;
  if (c0 < C) {
    long r = r0 + rl;
    for (; r + 96 < r1; r += 128) {         // four independent 16-byte loads in flight per thread
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) ld8<T>(x + (r + 32 * u) * C + c0, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += v[u][j]; q[j] = fmaf(v[u][j], v[u][j], q[j]); }
      cnt += 4.f;
    }
    for (; r < r1; r += 32) {
      float v[8];
      ld8<T>(x + r * C + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
      cnt += 1.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mean = cnt > 0.f ? s[j] / cnt : 0.f;
    red[rl][ct * 8 + j][0] = cnt;
    red[rl][ct * 8 + j][1] = mean;
    red[rl][ct * 8 + j][2] = cnt > 0.f ? fmaxf(q[j] - s[j] * mean, 0.f) : 0.f;
  }
  __syncthreads();
  // 4 threads per channel merge 8 row lanes each, then a 2-step shuffle merge
  const int ch = threadIdx.x >> 2, sub = threadIdx.x & 3;
  float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float* e = red[sub * 8 + k][ch];
    chan_merge(n, mean, m2, e[0], e[1], e[2]);
  }
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o), qb = __shfl_xor_sync(0xffffffffu, m2, o);
    chan_merge(n, mean, m2, nb, mb, qb);
  }
  if (sub == 0 && blockIdx.y * kBnCh + ch < C) {
    float* o = partial + ((static_cast<long>(blockIdx.y) * gridDim.x + blockIdx.x) * kBnCh + ch) * 3;
    o[0] = n; o[1] = mean; o[2] = m2;
  }
}

// One CTA per 64-channel block, 1024 threads = 64 channels x 16 part lanes, coalesced reads of the (L2-resident) partial
// table.  The partials are combined in two division-free passes instead of a chain of Chan updates (whose two divisions
// per merge made this kernel ~12 us on its single SM):  mean = sum n_i mean_i / N,  M2 = sum (M2_i + n_i (mean_i - mean)^2).
// Thread (channel, 0) produces scale / shift and updates the running statistics (momentum, unbiased variance: PyTorch).
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int nparts, int C,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* running_mean, float* running_var, float momentum, float eps,
                                                           int update_running, float* __restrict__ scale_shift /* [2][C] */) {
  __shared__ float red[16][kBnCh][2];
  __shared__ float s_mean[kBnCh], s_n[kBnCh];
  pdl_sync();
  const int ch = threadIdx.x & (kBnCh - 1), pl = threadIdx.x >> 6;
  const int c = blockIdx.x * kBnCh + ch;
  const float* base = partial + static_cast<long>(blockIdx.x) * nparts * kBnCh * 3;
  float n = 0.f, sm = 0.f;
  for (int p = pl; p < nparts; p += 16) {
    const float* e = base + (static_cast<long>(p) * kBnCh + ch) * 3;
    n += e[0];
    sm = fmaf(e[0], e[1], sm);
  }
  red[pl][ch][0] = n; red[pl][ch][1] = sm;
  __syncthreads();
  if (pl == 0) {
    float N = 0.f, S = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { N += red[k][ch][0]; S += red[k][ch][1]; }
    s_n[ch] = N;
    s_mean[ch] = N > 0.f ? S / N : 0.f;
  }
  __syncthreads();
  const float mean = s_mean[ch];
  float m2 = 0.f;
  for (int p = pl; p < nparts; p += 16) {
    const float* e = base + (static_cast<long>(p) * kBnCh + ch) * 3;
    const float d = e[1] - mean;
    m2 += fmaf(e[0] * d, d, e[2]);
  }
  red[pl][ch][0] = m2;
  __syncthreads();
  if (pl != 0 || c >= C) return;
  m2 = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) m2 += red[k][ch][0];
  n = s_n[ch];
  const float var = m2 / n;                      // biased: what normalisation uses
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[c] * invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = beta[c] - mean * sc;
  if (update_running) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (n > 1.f ? m2 / (n - 1.f) : var);
  }
}

__global__ void bn_eval_scale_kernel(int C, const float* gamma, const float* beta, const float* running_mean,
                                     const float* running_var, float eps, float* scale_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale_shift[c] = sc;
  scale_shift[C + c] = beta[c] - running_mean[c] * sc;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ resid, T* __restrict__ y,
                                                       long n8, int C, const float* __restrict__ scale_shift, int relu) {
  pdl_sync();
  const int c8 = C >> 3;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c8) * 8;
    float v[8], sc[8], sh[8];
    ld8<T>(x + i * 8, v);
    ld8<float>(scale_shift + ch, sc);
    ld8<float>(scale_shift + C + ch, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] * sc[j] + sh[j];
    if (resid) {
      float r[8];
      ld8<T>(resid + i * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    st8<T>(y + i * 8, v);
  }
}

}  // namespace

// workspace: [channel blocks][nparts][64][3] partials followed by [2 * C] scale/shift; nparts = bn_num_parts(rows, C)
int bn_num_parts(long rows, int C) {
  const int sms = device_sm_count();
  const int cblocks = (C + kBnCh - 1) / kBnCh;
  long parts = (static_cast<long>(sms) * 4 + cblocks - 1) / cblocks;       // ~4 CTAs per SM over the whole grid
  const long max_parts = (rows + 32 * 8 - 1) / (32 * 8);                    // at least 8 rows per thread
  if (parts > max_parts) parts = max_parts;
  if (parts < 1) parts = 1;
  return static_cast<int>(parts);
}
static long bn_partial_floats(long rows, int C) {
  return static_cast<long>((C + kBnCh - 1) / kBnCh) * bn_num_parts(rows, C) * kBnCh * 3;
}
long bn_workspace_floats(long rows, int C) { return bn_partial_floats(rows, C) + 2L * C; }

static int bn_launch_stats(const void* x, long rows, int C, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, float momentum, float eps, float* workspace, float* scale_shift, int f32,
                           cudaStream_t s) {
  const int nparts = bn_num_parts(rows, C);
  const int cblocks = (C + kBnCh - 1) / kBnCh;
  const long rows_per_cta = (rows + nparts - 1) / nparts;
  dim3 grid(nparts, cblocks);
  if (f32) MV_CUDA_CHECK(launch_pdl(bn_stats_kernel<float>, grid, dim3(256), 0, s, static_cast<const float*>(x), rows, C, rows_per_cta, workspace));
  else MV_CUDA_CHECK(launch_pdl(bn_stats_kernel<bf16>, grid, dim3(256), 0, s, static_cast<const bf16*>(x), rows, C, rows_per_cta, workspace));
  MV_LAUNCH_CHECK();
  MV_CUDA_CHECK(launch_pdl(bn_finalize_kernel, dim3(cblocks), dim3(1024), 0, s, static_cast<const float*>(workspace), nparts, C, gamma, beta,
                           running_mean, running_var, momentum, eps, 1, scale_shift));
  MV_LAUNCH_CHECK();
  return 0;
}

int bn_forward(const void* x, const void* resid, void* y, long rows, int C, const float* gamma, const float* beta,
               float* running_mean, float* running_var, float momentum, float eps, int training, int relu, float* workspace,
               long ws_floats, int f32, cudaStream_t s) {
  MV_REQUIRE(x && y && gamma && beta && running_mean && running_var && workspace, "bn_forward: null argument");
  MV_REQUIRE(C % 8 == 0 && C >= 8 && rows > 0, "bn_forward: C must be a multiple of 8 (got %d)", C);
  MV_REQUIRE(ws_floats >= bn_workspace_floats(rows, C), "bn_forward: workspace too small");
  float* scale_shift = workspace + bn_partial_floats(rows, C);
  if (training) {
    if (bn_launch_stats(x, rows, C, gamma, beta, running_mean, running_var, momentum, eps, workspace, scale_shift, f32, s)) return -2;
  } else {
    bn_eval_scale_kernel<<<(C + 127) / 128, 128, 0, s>>>(C, gamma, beta, running_mean, running_var, eps, scale_shift);
    MV_LAUNCH_CHECK();
  }
  const long n8 = rows * (C >> 3);
  long blocks = (n8 + 255) / 256;
  const long cap = static_cast<long>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  const dim3 agrid(static_cast<unsigned>(blocks));
  if (f32) MV_CUDA_CHECK(launch_pdl(bn_apply_kernel<float>, agrid, dim3(256), 0, s, static_cast<const float*>(x), static_cast<const float*>(resid),
                                    static_cast<float*>(y), n8, C, static_cast<const float*>(scale_shift), relu));
  else MV_CUDA_CHECK(launch_pdl(bn_apply_kernel<bf16>, agrid, dim3(256), 0, s, static_cast<const bf16*>(x), static_cast<const bf16*>(resid),
                                static_cast<bf16*>(y), n8, C, static_cast<const float*>(scale_shift), relu));
  MV_LAUNCH_CHECK();
  return 0;
}

}  // namespace mv

// ---- uint8 CXR -> normalised channels-last activation ------------------------------------------------------------------
// Fuses data/helper.py's ToTensor (u8 -> [0,1]) + Normalize(mean, std) with the NCHW -> NHWC layout change the cuDNN
// convolutions want, so a step ships 0.75 MB per image over PCIe instead of 3 MB of fp32.
namespace mv {
namespace {
template <typename T>
__global__ void __launch_bounds__(256) normalize_u8_kernel(const unsigned char* __restrict__ src, T* __restrict__ dst, long pixels,
                                                           long hw, int cpad, float m0, float m1, float m2, float s0, float s1, float s2) {
  // src: [B, 3, H, W] u8 ; dst: [B, H, W, cpad] (cpad >= 3; extra channels are zero)
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = i / hw, p = i - b * hw;
    const unsigned char* s = src + b * 3 * hw + p;
    const float r = (static_cast<float>(s[0]) * (1.f / 255.f) - m0) * s0;
    const float g = (static_cast<float>(s[hw]) * (1.f / 255.f) - m1) * s1;
    const float bl = (static_cast<float>(s[2 * hw]) * (1.f / 255.f) - m2) * s2;
    T* d = dst + i * cpad;
    d[0] = from_f32<T>(r); d[1] = from_f32<T>(g); d[2] = from_f32<T>(bl);
    for (int c = 3; c < cpad; ++c) d[c] = from_f32<T>(0.f);     // zero channels: lets cuDNN run the stem as a tensor-op conv
  }
}

// Space-to-depth form of the same transform, for the ResNet stem: the 7x7 / stride-2 / pad-3 convolution over 3 channels is
// run as a 4x4 / stride-1 convolution over 2x2 pixel blocks (12 channels, zero-padded to 16) — a shape cuDNN's tensor-core
// implicit GEMM handles well, where the 3-channel strided form runs at 0.39 TB/s (profiles/r01_launches_step_b64_summary.txt).
//   dst: [B, H/2 + 3, W/2 + 3, 16] channels-last, block (by, bx) at row by + 2, column bx + 2 (two zero rows / columns in
//        front, one behind: output pixel oy reads input rows 2oy - 4 .. 2oy + 3, i.e. blocks oy - 2 .. oy + 1);
//        channel = c * 4 + dy * 2 + dx (torch.nn.functional.pixel_unshuffle order), channels 12 .. 15 are zero.
template <typename T>
__global__ void __launch_bounds__(256) normalize_u8_s2d_kernel(const unsigned char* __restrict__ src, T* __restrict__ dst, int B, int H,
                                                               int W, float m0, float m1, float m2, float s0, float s1, float s2) {
  const int Hb = H >> 1, Wb = W >> 1, Hp = Hb + 3, Wp = Wb + 3;
  const long total = static_cast<long>(B) * Hp * Wp;
  const float mean[3] = {m0, m1, m2}, istd[3] = {s0, s1, s2};
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % Wp), py = static_cast<int>((i / Wp) % Hp), b = static_cast<int>(i / (static_cast<long>(Wp) * Hp));
    const int by = py - 2, bx = px - 2;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;
    if (by >= 0 && by < Hb && bx >= 0 && bx < Wb) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const unsigned char* pl = src + (static_cast<long>(b) * 3 + c) * H * W + static_cast<long>(2 * by) * W + 2 * bx;
        const uchar2 r0 = *reinterpret_cast<const uchar2*>(pl), r1 = *reinterpret_cast<const uchar2*>(pl + W);
        v[c * 4 + 0] = (static_cast<float>(r0.x) * (1.f / 255.f) - mean[c]) * istd[c];
        v[c * 4 + 1] = (static_cast<float>(r0.y) * (1.f / 255.f) - mean[c]) * istd[c];
        v[c * 4 + 2] = (static_cast<float>(r1.x) * (1.f / 255.f) - mean[c]) * istd[c];
        v[c * 4 + 3] = (static_cast<float>(r1.y) * (1.f / 255.f) - mean[c]) * istd[c];
      }
    }
    T* d = dst + i * 16;
    float lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { lo[k] = v[k]; hi[k] = v[8 + k]; }
    st8<T>(d, lo);
    st8<T>(d + 8, hi);
  }
}

// Stem tail fused in one pass: y = maxpool3x3/s2/p1( relu( x * scale + shift ) ) on channels-last [B, H, W, C] ->
// [B, H/2, W/2, C].  Replaces BatchNorm-apply + ReLU + nn.MaxPool2d (models/image.py:50-56, torchvision stem).
template <typename T>
__global__ void __launch_bounds__(256) bn_relu_maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C,
                                                              const float* __restrict__ scale_shift) {
  const int c8 = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long total = static_cast<long>(B) * Ho * Wo * c8;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c8) * 8;
    long t = i / c8;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    float sc[8], sh[8], best[8];
    ld8<float>(scale_shift + ch, sc);
    ld8<float>(scale_shift + C + ch, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = 0.f;               // relu output is >= 0 and every window holds >= 1 valid pixel
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int hi = 2 * ho + dy;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int wi = 2 * wo + dx;
        if (wi < 0 || wi >= W) continue;
        float v[8];
        ld8<T>(x + ((static_cast<long>(b) * H + hi) * W + wi) * C + ch, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], fmaf(v[j], sc[j], sh[j]));
      }
    }
    st8<T>(y + i * 8, best);
  }
}
}  // namespace

int bn_relu_maxpool(const void* x, void* y, int B, int H, int W, int C, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int training, float* workspace, long ws_floats, int f32,
                    cudaStream_t s) {
  MV_REQUIRE(x && y && gamma && beta && running_mean && running_var && workspace, "bn_relu_maxpool: null argument");
  MV_REQUIRE(C % 8 == 0 && C <= 2048 && H % 2 == 0 && W % 2 == 0, "bn_relu_maxpool: need C %% 8 == 0, even H and W");
  const long rows = static_cast<long>(B) * H * W;
  MV_REQUIRE(ws_floats >= bn_workspace_floats(rows, C), "bn_relu_maxpool: workspace too small");
  float* scale_shift = workspace + bn_partial_floats(rows, C);
  if (training) {
    if (bn_launch_stats(x, rows, C, gamma, beta, running_mean, running_var, momentum, eps, workspace, scale_shift, f32, s)) return -2;
  } else {
    bn_eval_scale_kernel<<<(C + 127) / 128, 128, 0, s>>>(C, gamma, beta, running_mean, running_var, eps, scale_shift);
    MV_LAUNCH_CHECK();
  }
  const long total = static_cast<long>(B) * (H / 2) * (W / 2) * (C / 8);
  long blocks = (total + 255) / 256;
  const long cap = static_cast<long>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (f32) bn_relu_maxpool_kernel<float><<<static_cast<int>(blocks), 256, 0, s>>>(static_cast<const float*>(x), static_cast<float*>(y), B, H, W, C, scale_shift);
  else bn_relu_maxpool_kernel<bf16><<<static_cast<int>(blocks), 256, 0, s>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), B, H, W, C, scale_shift);
  MV_LAUNCH_CHECK();
  return 0;
}

int normalize_u8_s2d(const unsigned char* src, void* dst, int B, int H, int W, const float mean[3], const float stdv[3], int f32,
                     cudaStream_t s) {
  MV_REQUIRE(src && dst && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "normalize_u8_s2d: need even H and W");
  const long total = static_cast<long>(B) * (H / 2 + 3) * (W / 2 + 3);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (f32) normalize_u8_s2d_kernel<float><<<static_cast<int>(blocks), 256, 0, s>>>(src, static_cast<float*>(dst), B, H, W, mean[0], mean[1], mean[2],
                                                                                  1.f / stdv[0], 1.f / stdv[1], 1.f / stdv[2]);
  else normalize_u8_s2d_kernel<bf16><<<static_cast<int>(blocks), 256, 0, s>>>(src, static_cast<bf16*>(dst), B, H, W, mean[0], mean[1], mean[2],
                                                                              1.f / stdv[0], 1.f / stdv[1], 1.f / stdv[2]);
  MV_LAUNCH_CHECK();
  return 0;
}

int normalize_u8(const unsigned char* src, void* dst, long B, long hw, int cpad, const float mean[3], const float stdv[3], int f32,
                 cudaStream_t s) {
  MV_REQUIRE(src && dst && B > 0 && hw > 0 && cpad >= 3 && cpad <= 8, "normalize_u8: bad arguments");
  const long pixels = B * hw;
  long blocks = (pixels + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (f32) normalize_u8_kernel<float><<<static_cast<int>(blocks), 256, 0, s>>>(src, static_cast<float*>(dst), pixels, hw, cpad, mean[0], mean[1], mean[2],
                                                                              1.f / stdv[0], 1.f / stdv[1], 1.f / stdv[2]);
  else normalize_u8_kernel<bf16><<<static_cast<int>(blocks), 256, 0, s>>>(src, static_cast<bf16*>(dst), pixels, hw, cpad, mean[0], mean[1], mean[2],
                                                                          1.f / stdv[0], 1.f / stdv[1], 1.f / stdv[2]);
  MV_LAUNCH_CHECK();
  return 0;
}
}  // namespace mv
