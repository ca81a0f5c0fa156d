// attention_simt.cu — fp32 check-mode masked attention (forward + backward) on CUDA cores.  Same semantics, mask
// predicate and dropout indexing as the tcgen05 kernels; used for the 1e-4 parity gate against the CPU oracle.
// Reference arithmetic: softmax(QK^T/sqrt(64) + (1-m)*-1e4) -> dropout -> .V  (upstream BertSelfAttention; in-tree twin
// Downstream_task/report_generation_and_vqa/sc/pytorch_pretrained_bert/model.py:301-320).
#include "attn_common.cuh"
#include "kernels.h"

namespace mv {
namespace {

constexpr int D = 64;

__device__ __forceinline__ float dot64(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < D; i += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + i), y = *reinterpret_cast<const float4*>(b + i);
    s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  return s;
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = scratch[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, scratch[w]) : r + scratch[w];
  return r;
}

__global__ void __launch_bounds__(128) attn_fwd_simt_kernel(const AttnArgs a) {
  extern __shared__ float sm[];  // [L] probabilities
  __shared__ __align__(16) float sq[D];
  __shared__ float scratch[4];
  const int q = blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int H = a.nh * D, L = a.L;
  const float* base = static_cast<const float*>(a.qkv) + static_cast<long>(b) * L * 3 * H;
  const int mode = a.mode[b], tl = a.t_len[b];
  if (tid < D) sq[tid] = base[static_cast<long>(q) * 3 * H + h * D + tid];
  __syncthreads();
  const float scale = 0.125f;
  float mx = -INFINITY;
  for (int k = tid; k < L; k += 128) {
    float s = -INFINITY;
    if (mask_allowed(mode, q, k, a.A, tl)) s = dot64(sq, base + static_cast<long>(k) * 3 * H + H + h * D) * scale;
    sm[k] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce(mx, true, scratch);
  float sum = 0.f;
  for (int k = tid; k < L; k += 128) {
    const float p = sm[k] == -INFINITY ? 0.f : __expf(sm[k] - mx);
    sum += p;
    float pd = p;
    if (a.drop_on) pd = attn_keep(a.drop, a.drop_site, b, a.nh, h, L, q, k) ? p * a.drop.scale : 0.f;
    sm[k] = pd;
  }
  sum = block_reduce(sum, false, scratch);
  __syncthreads();
  if (tid < D) {
    float o = 0.f;
    for (int k = 0; k < L; ++k) o += sm[k] * base[static_cast<long>(k) * 3 * H + 2 * H + h * D + tid];
    static_cast<float*>(a.ctx)[(static_cast<long>(b) * L + q) * H + h * D + tid] = o / sum;
  }
  if (tid == 0) a.lse[(static_cast<long>(b) * a.nh + h) * L + q] = mx + logf(sum);
}

// dQ (and delta) — one CTA per query row
__global__ void __launch_bounds__(128) attn_bwd_dq_simt_kernel(const AttnArgs a) {
  extern __shared__ float sm[];  // [L] dS
  __shared__ __align__(16) float sq[D];
  __shared__ __align__(16) float sdo[D];
  __shared__ float scratch[4];
  const int q = blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int H = a.nh * D, L = a.L;
  const float* base = static_cast<const float*>(a.qkv) + static_cast<long>(b) * L * 3 * H;
  const long orow = (static_cast<long>(b) * L + q) * H + h * D;
  const int mode = a.mode[b], tl = a.t_len[b];
  float dpart = 0.f;
  if (tid < D) {
    sq[tid] = base[static_cast<long>(q) * 3 * H + h * D + tid];
    sdo[tid] = static_cast<const float*>(a.dctx)[orow + tid];
    dpart = sdo[tid] * static_cast<const float*>(a.ctx)[orow + tid];
  }
  const float delta = block_reduce(dpart, false, scratch);
  const float lse = a.lse[(static_cast<long>(b) * a.nh + h) * L + q];
  if (tid == 0) a.delta[(static_cast<long>(b) * a.nh + h) * L + q] = delta;
  const float scale = 0.125f;
  for (int k = tid; k < L; k += 128) {
    float ds = 0.f;
    if (mask_allowed(mode, q, k, a.A, tl)) {
      const float* kr = base + static_cast<long>(k) * 3 * H + H + h * D;
      const float p = __expf(dot64(sq, kr) * scale - lse);
      float dp = dot64(sdo, kr + H);
      if (a.drop_on) dp = attn_keep(a.drop, a.drop_site, b, a.nh, h, L, q, k) ? dp * a.drop.scale : 0.f;
      ds = p * (dp - delta) * scale;
    }
    sm[k] = ds;
  }
  __syncthreads();
  if (tid < D) {
    float acc = 0.f;
    for (int k = 0; k < L; ++k) acc += sm[k] * base[static_cast<long>(k) * 3 * H + H + h * D + tid];
    static_cast<float*>(a.dqkv)[(static_cast<long>(b) * L + q) * 3 * H + h * D + tid] = acc;
  }
}

// dK, dV — one CTA per key row
__global__ void __launch_bounds__(128) attn_bwd_dkv_simt_kernel(const AttnArgs a) {
  extern __shared__ float sm[];  // [2][L]: P_drop, dS
  __shared__ __align__(16) float sk[D];
  __shared__ __align__(16) float sv[D];
  const int k = blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int H = a.nh * D, L = a.L;
  const float* base = static_cast<const float*>(a.qkv) + static_cast<long>(b) * L * 3 * H;
  const float* dctx = static_cast<const float*>(a.dctx) + static_cast<long>(b) * L * H;
  const int mode = a.mode[b], tl = a.t_len[b];
  if (tid < D) {
    sk[tid] = base[static_cast<long>(k) * 3 * H + H + h * D + tid];
    sv[tid] = base[static_cast<long>(k) * 3 * H + 2 * H + h * D + tid];
  }
  __syncthreads();
  const float scale = 0.125f;
  float* sp = sm;
  float* sds = sm + L;
  for (int q = tid; q < L; q += 128) {
    float pd = 0.f, ds = 0.f;
    if (mask_allowed(mode, q, k, a.A, tl)) {
      const long st = (static_cast<long>(b) * a.nh + h) * L + q;
      const float p = __expf(dot64(base + static_cast<long>(q) * 3 * H + h * D, sk) * scale - a.lse[st]);
      float dp = dot64(dctx + static_cast<long>(q) * H + h * D, sv);
      pd = p;
      if (a.drop_on) {
        const bool keep = attn_keep(a.drop, a.drop_site, b, a.nh, h, L, q, k);
        dp = keep ? dp * a.drop.scale : 0.f;
        pd = keep ? p * a.drop.scale : 0.f;
      }
      ds = p * (dp - a.delta[st]) * scale;
    }
    sp[q] = pd;
    sds[q] = ds;
  }
  __syncthreads();
  const int d = tid & 63;
  float acc = 0.f;
  if (tid < 64) {
    for (int q = 0; q < L; ++q) acc += sp[q] * dctx[static_cast<long>(q) * H + h * D + d];
    static_cast<float*>(a.dqkv)[(static_cast<long>(b) * L + k) * 3 * H + 2 * H + h * D + d] = acc;
  } else {
    for (int q = 0; q < L; ++q) acc += sds[q] * base[static_cast<long>(q) * 3 * H + h * D + d];
    static_cast<float*>(a.dqkv)[(static_cast<long>(b) * L + k) * 3 * H + H + h * D + d] = acc;
  }
}

}  // namespace

int attention_fwd_simt(const AttnArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.qkv && a.ctx && a.lse && a.mode && a.t_len, "attention_fwd: null argument");
  dim3 grid(a.L, a.nh, a.B);
  attn_fwd_simt_kernel<<<grid, 128, a.L * sizeof(float), s>>>(a);
  MV_LAUNCH_CHECK();
  return 0;
}

int attention_bwd_simt(const AttnArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.qkv && a.ctx && a.lse && a.dctx && a.dqkv && a.delta, "attention_bwd: null argument");
  dim3 grid(a.L, a.nh, a.B);
  attn_bwd_dq_simt_kernel<<<grid, 128, a.L * sizeof(float), s>>>(a);
  MV_LAUNCH_CHECK();
  attn_bwd_dkv_simt_kernel<<<grid, 128, 2 * a.L * sizeof(float), s>>>(a);
  MV_LAUNCH_CHECK();
  return 0;
}

}  // namespace mv
