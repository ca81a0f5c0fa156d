// loss_optim.cu — MLM softmax-cross-entropy (forward + gradient in one pass over the labelled rows' logits),
// ITM head (Linear(768,2) + CE + backward through tanh), fused multi-tensor AdamW with bf16 shadow refresh.
#include "kernels.h"

namespace mv {
namespace {

// One CTA per labelled row.  Pass 1: running (max, first argmax); pass 2: sum of exp; pass 3: gradient.
// Reference: nn.CrossEntropyLoss(ignore_index=-100) over logits.transpose(1,2) (models/train_origin.py:62,120) —
// mean over labelled tokens == sum_i (lse_i - z_i[y_i]) * gscale with gscale = 1 / n_labelled; and the MLM accuracy
// argmax of models/train_origin.py:138-146 (torch.argmax returns the first maximal index).
template <typename T>
__global__ void __launch_bounds__(256) mlm_ce_kernel(const CeArgs a) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ float s_bcast[2];
  __shared__ int s_arg;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* z = a.logits + static_cast<long>(row) * a.ldv;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int c = tid; c < a.V; c += 256) {
    const float v = z[c];
    if (v > m) { m = v; mi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  if (lane == 0) { s_val[warp] = m; s_idx[warp] = mi; }
  __syncthreads();
  if (tid == 0) {
    float bm = s_val[0];
    int bi = s_idx[0];
    for (int w = 1; w < 8; ++w)
      if (s_val[w] > bm || (s_val[w] == bm && s_idx[w] < bi)) { bm = s_val[w]; bi = s_idx[w]; }
    s_bcast[0] = bm;
    s_arg = bi;
  }
  __syncthreads();
  const float mx = s_bcast[0];
  float sum = 0.f;
  for (int c = tid; c < a.V; c += 256) sum += __expf(z[c] - mx);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) s_val[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_val[w];
    s_bcast[1] = t;
  }
  __syncthreads();
  const float total = s_bcast[1];
  const float inv = 1.0f / total;
  int label = static_cast<int>(a.labels[row]);
  if (label < 0 || label >= a.V) {       // PyTorch's CE raises for a target outside [0, V): clamp and report
    if (tid == 0 && a.err) atomicOr(a.err, 8);
    label = 0;
  }
  const float rw = a.row_weight ? a.row_weight[row] : 1.f;       // fine-tune masked_weights (model.py:998-1005)
  if (a.dlogits) {
    T* d = static_cast<T*>(a.dlogits) + static_cast<long>(row) * a.ldv;
    const float gs = (a.count_dev ? 1.f / *a.count_dev : a.gscale) * rw;
    for (int c = tid; c < a.ldv; c += 256) {
      float g = 0.f;
      if (c < a.V) g = (__expf(z[c] - mx) * inv - (c == label ? 1.f : 0.f)) * gs;
      d[c] = from_f32<T>(g);
    }
  }
  if (tid == 0) {
    const float lse = mx + logf(total);
    atomicAdd(a.loss_sum, (lse - z[label]) * rw);
    if (s_arg == label) atomicAdd(a.correct, 1);
    if (a.row_lse) a.row_lse[row] = lse;
    if (a.row_argmax) a.row_argmax[row] = s_arg;
    if (a.row_loss) a.row_loss[row] = lse - z[label];
  }
}

// One CTA; B <= 1024 samples.  Dynamic smem: loss[B], wsum[B], keep[B].
__global__ void __launch_bounds__(1024) drop_worst_kernel(const DropWorstArgs a) {
  extern __shared__ float dw_smem[];
  float* s_loss = dw_smem;
  float* s_w = dw_smem + a.B;
  int* s_keep = reinterpret_cast<int*>(dw_smem + 2 * a.B);
  __shared__ float s_acc[2];
  const int tid = threadIdx.x;
  for (int b = tid; b < a.B; b += blockDim.x) { s_loss[b] = 0.f; s_w[b] = 0.f; }
  if (tid < 2) s_acc[tid] = 0.f;
  __syncthreads();
  for (int i = tid; i < a.n; i += blockDim.x) {
    const int b = static_cast<int>(a.rows[i] / a.L);
    const float w = a.row_weight ? a.row_weight[i] : 1.f;
    atomicAdd(&s_loss[b], a.row_loss[i] * w);
    atomicAdd(&s_w[b], w);
  }
  __syncthreads();
  for (int b = tid; b < a.B; b += blockDim.x) {
    const float mine = s_loss[b];
    int rank = 0;
    for (int j = 0; j < a.B; ++j) {
      const float o = s_loss[j];
      rank += (o < mine || (o == mine && j < b)) ? 1 : 0;
    }
    const int k = rank < a.keep ? 1 : 0;
    s_keep[b] = k;
    if (k) { atomicAdd(&s_acc[0], mine); atomicAdd(&s_acc[1], s_w[b]); }
  }
  __syncthreads();
  const float inv = 1.f / (s_acc[1] + 1e-5f);
  for (int i = tid; i < a.n; i += blockDim.x) {
    const int b = static_cast<int>(a.rows[i] / a.L);
    const float w = a.row_weight ? a.row_weight[i] : 1.f;
    a.row_scale[i] = s_keep[b] ? w * inv : 0.f;
  }
  if (tid == 0) atomicAdd(a.loss_sum, s_acc[0] * inv);
}

// One CTA per sample.  logits = pooled . W^T + b (models/cxrbert_origin.py:164-173); CE mean over the batch
// (models/train_origin.py:63,123); d_pre = (dlogits . W) * (1 - pooled^2) is the gradient entering the pooler GEMM.
template <typename T>
__global__ void __launch_bounds__(256) itm_kernel(const ItmArgs a) {
  __shared__ float s_red[2][8];
  __shared__ float s_dl[2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* p = static_cast<const T*>(a.pooled) + static_cast<long>(b) * a.H;
  float a0 = 0.f, a1 = 0.f;
  for (int j = tid; j < a.H; j += 256) {
    const float x = to_f32<T>(p[j]);
    a0 += x * a.w[j];
    a1 += x * a.w[a.H + j];
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (lane == 0) { s_red[0][warp] = a0; s_red[1][warp] = a1; }
  __syncthreads();
  if (tid == 0) {
    float z0 = a.b[0], z1 = a.b[1];
    for (int w = 0; w < 8; ++w) { z0 += s_red[0][w]; z1 += s_red[1][w]; }
    a.logits[2 * b] = z0;
    a.logits[2 * b + 1] = z1;
    if (a.ext_dlogits != nullptr) {                              // gradient handed in by the caller (autograd path)
      const float d0 = a.ext_dlogits[2 * b], d1 = a.ext_dlogits[2 * b + 1];
      s_dl[0] = d0; s_dl[1] = d1;
      if (a.d_pre) { atomicAdd(a.db, d0); atomicAdd(a.db + 1, d1); }
    } else if (a.labels == nullptr) { s_dl[0] = 0.f; s_dl[1] = 0.f; }   // forward-only: logits are all that is asked
    else {
    const int y = static_cast<int>(a.labels[b]);
    const float mx = fmaxf(z0, z1);
    const float e0 = __expf(z0 - mx), e1 = __expf(z1 - mx);
    const float lse = mx + logf(e0 + e1);
    atomicAdd(a.loss_sum, lse - (y == 0 ? z0 : z1));
    const int pred = z1 > z0 ? 1 : 0;  // argmax returns the first maximal index on ties
    if (pred == y) atomicAdd(a.correct, 1);
    const float gsc = a.count_dev ? 1.f / *a.count_dev : a.gscale;
    const float d0 = (e0 / (e0 + e1) - (y == 0 ? 1.f : 0.f)) * gsc;
    const float d1 = (e1 / (e0 + e1) - (y == 1 ? 1.f : 0.f)) * gsc;
    s_dl[0] = d0;
    s_dl[1] = d1;
    if (a.d_pre) { atomicAdd(a.db, d0); atomicAdd(a.db + 1, d1); }
    }
  }
  __syncthreads();
  if (a.d_pre) {
    const float d0 = s_dl[0], d1 = s_dl[1];
    T* dp = static_cast<T*>(a.d_pre) + static_cast<long>(b) * a.H;
    for (int j = tid; j < a.H; j += 256) {
      const float x = to_f32<T>(p[j]);
      atomicAdd(a.dw + j, d0 * x);
      atomicAdd(a.dw + a.H + j, d1 * x);
      dp[j] = from_f32<T>((d0 * a.w[j] + d1 * a.w[a.H + j]) * (a.no_tanh ? 1.f : (1.f - x * x)));
    }
  }
}

// HF-3.x AdamW (correct_bias=True): p -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps); decoupled decay after.
__global__ void __launch_bounds__(256) adamw_kernel(const AdamArgs a, float step_size) {
  const long n4 = a.n >> 2;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    float4 g = reinterpret_cast<float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float* pp = &p.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * a.grad_scale;
      mp[j] = mp[j] * a.beta1 + gj * (1.f - a.beta1);
      vp[j] = vp[j] * a.beta2 + gj * gj * (1.f - a.beta2);
      pp[j] = pp[j] - step_size * (mp[j] / (sqrtf(vp[j]) + a.eps));
      if (a.weight_decay > 0.f) pp[j] = pp[j] - a.lr * a.weight_decay * pp[j];
    }
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.zero_grad) reinterpret_cast<float4*>(a.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.shadow) {
      uint2 s;
      s.x = pack_bf16x2(p.x, p.y);
      s.y = pack_bf16x2(p.z, p.w);
      reinterpret_cast<uint2*>(a.shadow)[i] = s;
    }
  }
}

// ---- BertAdam (fine-tune optimizer, Downstream_task/report_generation_and_vqa/sc/pytorch_pretrained_bert/optimization.py:112-182)
// The arena is cut into chunks that never straddle a parameter tensor; pass 1 accumulates each tensor's squared gradient
// norm, pass 2 applies clip_grad_norm_(p, max_norm) PER TENSOR (:146-147), Adam moments WITHOUT bias correction (:151-153,
// :178-181), decoupled weight decay added to the update (:162-163) and the scheduled learning rate (:165-172).
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, const AdamChunk* __restrict__ chunks,
                                                         float* __restrict__ sumsq) {
  const AdamChunk c = chunks[blockIdx.x];
  float acc = 0.f;
  for (int i = threadIdx.x; i < c.n; i += 256) { const float x = g[c.off + i]; acc += x * x; }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(sumsq + c.tensor, t);
  }
}

__global__ void __launch_bounds__(256) bert_adam_kernel(const BertAdamArgs a) {
  const AdamChunk c = a.chunks[blockIdx.x];
  float coef = 1.f;
  if (a.max_grad_norm > 0.f) {
    coef = a.max_grad_norm / (sqrtf(a.sumsq[c.tensor]) + 1e-6f);      // torch.nn.utils.clip_grad_norm_
    coef = fminf(coef, 1.f);
  }
  const float wd = c.decay ? a.weight_decay : 0.f;
  for (int i = threadIdx.x; i < c.n; i += 256) {
    const long j = c.off + i;
    const float g = a.g[j] * coef;
    const float m = a.m[j] * a.beta1 + (1.f - a.beta1) * g;
    const float v = a.v[j] * a.beta2 + (1.f - a.beta2) * g * g;
    float p = a.p[j];
    const float update = m / (sqrtf(v) + a.eps) + wd * p;
    p -= a.lr * update;
    a.p[j] = p; a.m[j] = m; a.v[j] = v;
    a.g[j] = 0.f;
    if (a.shadow) a.shadow[j] = __float2bfloat16_rn(p);
  }
}

// Retrieval score of a pair = softmax(itm_logits)[1] (Downstream_task/Retrieval/full_dset_retrieval.py:506-507), fp32
__global__ void __launch_bounds__(256) itm_match_prob_kernel(const float* __restrict__ logits, float* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float z0 = logits[2 * b], z1 = logits[2 * b + 1];
  const float mx = fmaxf(z0, z1);
  const float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
  out[b] = e1 / (e0 + e1);
}

}  // namespace

int bert_adam_step(const BertAdamArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.p && a.g && a.m && a.v && a.chunks && a.sumsq && a.n_chunks > 0 && a.n_tensors > 0, "bert_adam: null argument");
  MV_CUDA_CHECK(cudaMemsetAsync(a.sumsq, 0, sizeof(float) * a.n_tensors, s));
  grad_sumsq_kernel<<<a.n_chunks, 256, 0, s>>>(a.g, a.chunks, a.sumsq);
  MV_LAUNCH_CHECK();
  bert_adam_kernel<<<a.n_chunks, 256, 0, s>>>(a);
  MV_LAUNCH_CHECK();
  return 0;
}

int itm_match_prob(const float* logits, float* out, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  MV_REQUIRE(logits && out, "itm_match_prob: null argument");
  itm_match_prob_kernel<<<(B + 255) / 256, 256, 0, s>>>(logits, out, B);
  MV_LAUNCH_CHECK();
  return 0;
}

int mlm_ce_fwd_bwd(const CeArgs& a, int f32, cudaStream_t s) {
  if (a.n <= 0) return 0;
  MV_REQUIRE(a.logits && a.labels && a.loss_sum && a.correct, "mlm_ce: null argument");
  if (f32) mlm_ce_kernel<float><<<a.n, 256, 0, s>>>(a); else mlm_ce_kernel<bf16><<<a.n, 256, 0, s>>>(a);
  MV_LAUNCH_CHECK();
  return 0;
}

int drop_worst_select(const DropWorstArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.n > 0 && a.B > 0 && a.B <= 1024 && a.L > 0, "drop_worst: needs 0 < B <= 1024 and labelled rows");
  MV_REQUIRE(a.keep >= 1 && a.keep <= a.B, "drop_worst: int(B * (1 - ratio)) = %d samples kept of %d", a.keep, a.B);
  MV_REQUIRE(a.row_loss && a.rows && a.row_scale && a.loss_sum, "drop_worst: null argument");
  drop_worst_kernel<<<1, 1024, 3 * a.B * sizeof(float), s>>>(a);
  MV_LAUNCH_CHECK();
  return 0;
}

int itm_head_fwd_bwd(const ItmArgs& a, int f32, cudaStream_t s) {
  if (a.B <= 0) return 0;
  MV_REQUIRE(a.pooled && a.w && a.b && a.logits && a.loss_sum && a.correct, "itm: null argument");
  MV_REQUIRE(a.labels || a.ext_dlogits || !a.d_pre, "itm: backward needs labels or an external logit gradient");
  if (f32) itm_kernel<float><<<a.B, 256, 0, s>>>(a); else itm_kernel<bf16><<<a.B, 256, 0, s>>>(a);
  MV_LAUNCH_CHECK();
  return 0;
}

int adamw_step(const AdamArgs& a, cudaStream_t s) {
  MV_REQUIRE(a.n % 4 == 0, "adamw: arena length must be a multiple of 4");
  MV_REQUIRE(a.step >= 1, "adamw: step starts at 1");
  const double bc1 = 1.0 - pow(static_cast<double>(a.beta1), a.step);
  const double bc2 = 1.0 - pow(static_cast<double>(a.beta2), a.step);
  const float step_size = static_cast<float>(a.lr * sqrt(bc2) / bc1);
  adamw_kernel<<<148 * 8, 256, 0, s>>>(a, step_size);
  MV_LAUNCH_CHECK();
  return 0;
}

}  // namespace mv
