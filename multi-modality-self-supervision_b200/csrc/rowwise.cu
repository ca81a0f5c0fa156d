// rowwise.cu — HBM-bound fused kernels: joint embedding-sum + LayerNorm (+dropout), LayerNorm forward/backward with
// dropout-backward and bias/affine gradient reductions, row gather/scatter, column sums, GELU backward, casts, and the
// on-the-fly mask dump / classifier.  One warp owns one row; every global access is a 16-byte vector; rows are
// 8-element aligned (H % 8 == 0, H <= 1024).  Templated on the activation type (bf16 production, fp32 check mode).
#include "kernels.h"

namespace mv {
namespace {

constexpr int kMaxChunks = 4;  // 8-element chunks per lane: H <= 32 * 4 * 8 = 1024

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ void atomic_add8(float* p, const float (&v)[8]) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  atomicAdd(reinterpret_cast<float4*>(p + 4), make_float4(v[4], v[5], v[6], v[7]));
}

// mean / rstd of a row held as nchunk x 8 values per lane (two-pass, fp32)
__device__ __forceinline__ void row_stats(const float (&x)[kMaxChunks][8], int nch, int lane, int H, float eps,
                                          float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c)
    if (lane + 32 * c < nch)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += x[c][j];
  mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c)
    if (lane + 32 * c < nch)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = x[c][j] - mean; q += d * d; }
  rstd = rsqrtf(warp_sum(q) / H + eps);
}

__device__ __forceinline__ void apply_dropout8(float (&v)[8], const DropoutCfg& d, uint32_t site, long row, int H, int col) {
  const uint32_t keep = dropout_keep8(d, site, static_cast<uint64_t>(row) * H + col);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * d.scale : 0.f;
}

// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) embed_ln_fwd_kernel(const EmbedArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int rows = a.B * a.L;
  if (warp >= rows) return;
  const int b = warp / a.L, s = warp % a.L;
  const int H = a.H, nch = H >> 3;
  const float* wrow = nullptr;
  const T* prow = nullptr;
  int pos_id = 0, type_id = 0;
  // ids index the embedding tables: PyTorch raises IndexError when one is out of range; here it is clamped to row 0 and
  // reported through the step's error flags (a tokenizer vocabulary larger than config.vocab_size is the usual cause)
  auto checked = [&](long id, int hi, int flag) -> long {
    if (hi > 0 && (id < 0 || id >= hi)) { if (lane == 0 && a.err) atomicOr(a.err, flag); return 0; }
    return id;
  };
  if (s == 0) {
    wrow = a.word + checked(a.cls_tok[b], a.V, 1) * H;                        // [CLS]: position 0, prefix type
    type_id = a.prefix_type;
  } else if (s <= a.N) {
    prow = static_cast<const T*>(a.proj) + (static_cast<long>(b) * a.N + (s - 1)) * H;
    pos_id = static_cast<int>(checked(a.region_idx[s - 1], a.P, 4));          // grid index as position id
    type_id = a.prefix_type;
  } else if (s == a.N + 1) {
    wrow = a.word + checked(a.sep_tok[b], a.V, 1) * H;                        // [SEP]: position restarts at 0 (pre-training)
    pos_id = a.sep_pos;                                                       // or continues at A-1 (fine-tune model)
    type_id = a.prefix_type;
  } else {
    const int i = s - a.A;
    wrow = a.word + checked(a.input_ids[static_cast<long>(b) * a.T + i], a.V, 1) * H;
    pos_id = i;
    type_id = static_cast<int>(checked(a.segment[static_cast<long>(b) * a.T + i], a.TV, 2));
  }
  const float* posrow = a.pos + static_cast<long>(pos_id) * H;
  const float* typerow = a.type + static_cast<long>(type_id) * H;
  float x[kMaxChunks][8];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) {
      float w[8], p[8], t[8];
      if (wrow) load8<float>(wrow + ch * 8, w); else load8<T>(prow + ch * 8, w);
      load8<float>(posrow + ch * 8, p);
      load8<float>(typerow + ch * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[c][j] = w[j] + p[j] + t[j];
      store8<T>(static_cast<T*>(a.emb_sum) + static_cast<long>(warp) * H + ch * 8, x[c]);
    }
  }
  float mean, rstd;
  row_stats(x, nch, lane, H, a.eps, mean, rstd);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) {
      float g[8], bt[8], y[8];
      load8<float>(a.gamma + ch * 8, g);
      load8<float>(a.beta + ch * 8, bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (x[c][j] - mean) * rstd * g[j] + bt[j];
      if (a.drop_on) apply_dropout8(y, (s >= 1 && s <= a.N) ? a.drop_img : a.drop, a.drop_site, warp, H, ch * 8);
      store8<T>(static_cast<T*>(a.out) + static_cast<long>(warp) * H + ch * 8, y);
      if (a.out32) store8<float>(a.out32 + static_cast<long>(warp) * H + ch * 8, y);
    }
  }
}

// TX: dtype of the pre-LN sum (fp32 when the residual stream is kept in fp32), T: activation dtype of the
// output; y32 (optional): the same output unrounded, the fp32 residual operand of the next GEMM epilogue.
template <typename TX, typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TX* __restrict__ x, T* __restrict__ y, float* __restrict__ y32,
                                                     const float* gamma, const float* beta, int rows, int H, float eps,
                                                     int drop_on, uint32_t site, DropoutCfg drop) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nch = H >> 3;
  float v[kMaxChunks][8];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) load8<TX>(x + static_cast<long>(warp) * H + ch * 8, v[c]);
  }
  float mean, rstd;
  row_stats(v, nch, lane, H, eps, mean, rstd);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) {
      float g[8], bt[8], o[8];
      load8<float>(gamma + ch * 8, g);
      load8<float>(beta + ch * 8, bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[c][j] - mean) * rstd * g[j] + bt[j];
      if (drop_on) apply_dropout8(o, drop, site, warp, H, ch * 8);
      store8<T>(y + static_cast<long>(warp) * H + ch * 8, o);
      if (y32) store8<float>(y32 + static_cast<long>(warp) * H + ch * 8, o);
    }
  }
}

// LayerNorm backward.  8 warps / CTA, each warp strides over rows.  The column reductions (dgamma, dbeta, dbias) are
// accumulated in WARP-PRIVATE shared-memory rows (each lane owns its columns, so no synchronisation), which keeps the
// register budget low enough for two CTAs per SM; at the end the 8 rows are summed and one fp32 atomicAdd per column
// per CTA goes to the gradient arena.  Algorithmic traffic: read dy, x; write dx (+ dx_drop).
// A row costs ONE butterfly reduction: with x' = x - x0 (x0 = the row's first element, a cheap shift that keeps the
// single-pass variance well conditioned) the four sums  sum x', sum x'^2, sum dy*g, sum dy*g*x'  give mean, rstd and both
// projection terms  s1 = mean(dy*g),  s2 = mean(dy*g*xhat) = rstd * (sum dy*g*x' - mean' * sum dy*g) / H,
// so the dependency chain per row is load -> one 4-wide reduction -> store (it was four serial reductions).  The bf16
// instantiation also prefetches the next row's operands (raw 16-byte vectors) under the current row's arithmetic.
template <typename T>
struct Raw8;                                 // 8 consecutive elements as raw 16-byte vectors
template <>
struct Raw8<bf16> { uint4 u; };
template <>
struct Raw8<float> { float4 u[2]; };
__device__ __forceinline__ void raw_load8(Raw8<bf16>& r, const bf16* p) { r.u = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void raw_load8(Raw8<float>& r, const float* p) {
  r.u[0] = *reinterpret_cast<const float4*>(p);
  r.u[1] = *reinterpret_cast<const float4*>(p + 4);
}
template <typename TX, typename T, int NC>
struct RawRow { Raw8<TX> x[NC]; Raw8<T> d[NC]; };

template <typename TX, typename T, int NC>
__device__ __forceinline__ void raw_load(RawRow<TX, T, NC>& r, const TX* x, const T* dy, long row, int H, int nch, int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) {
      raw_load8(r.x[c], x + row * H + ch * 8);
      raw_load8(r.d[c], dy + row * H + ch * 8);
    }
  }
}
__device__ __forceinline__ void raw_unpack(const Raw8<bf16>& r, float (&v)[8]) {
  const uint4& u = r.u;
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void raw_unpack(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.u[0].x; v[1] = r.u[0].y; v[2] = r.u[0].z; v[3] = r.u[0].w; v[4] = r.u[1].x; v[5] = r.u[1].y; v[6] = r.u[1].z; v[7] = r.u[1].w;
}

template <typename TX, typename T, int NC>
__global__ void __launch_bounds__(256, 1) ln_bwd_kernel(const T* __restrict__ dy, const TX* __restrict__ x,
                                                        const float* __restrict__ gamma, T* __restrict__ dx,
                                                        T* __restrict__ dx_drop, float* dgamma, float* dbeta, float* dbias,
                                                        int rows, int H, float eps, int in_drop, int out_drop, uint32_t site,
                                                        DropoutCfg drop, LnAltDrop alt) {
  extern __shared__ float red[];  // [8 warps][3][H]
  constexpr bool kPrefetch = sizeof(T) == 2;      // production mode (bf16 gradients): prefetch the next row under the arithmetic
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = H >> 3;
  const float inv_h = 1.f / static_cast<float>(H);
  float* my = red + static_cast<size_t>(warp) * 3 * H;
  // column sums of this warp's rows live in registers (a lane owns the same 8-column chunks in every row); the r01 version
  // kept them in shared memory and paid 36 LDS/STS per row per lane, the kernel's top stall (short scoreboard, 34 %)
  float acc_g[NC][8], acc_b[NC][8], acc_bias[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_g[c][j] = acc_b[c][j] = acc_bias[c][j] = 0.f;
  pdl_sync();
  const long stride = static_cast<long>(gridDim.x) * 8;
  long row = static_cast<long>(blockIdx.x) * 8 + warp;
  RawRow<TX, T, NC> cur;
  if (kPrefetch && row < rows) raw_load(cur, x, dy, row, H, nch, lane);
  for (; row < rows; row += stride) {
    if (!kPrefetch) raw_load(cur, x, dy, row, H, nch, lane);
    float xv[NC][8], dv[NC][8];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        raw_unpack(cur.x[c], xv[c]);
        raw_unpack(cur.d[c], dv[c]);
        if (in_drop) {
          const int rr = alt.period > 0 ? static_cast<int>(row % alt.period) : -1;
          apply_dropout8(dv[c], (rr >= alt.lo && rr < alt.hi) ? alt.drop : drop, site, row, H, ch * 8);
        }
      }
    }
    if (kPrefetch && row + stride < rows) raw_load(cur, x, dy, row + stride, H, nch, lane);
    const float x0 = __shfl_sync(0xffffffffu, xv[0][0], 0);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        float g[8];
        load8<float>(gamma + ch * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xs = xv[c][j] - x0;
          const float dg = dv[c][j] * g[j];
          a0 += xs;
          a1 = fmaf(xs, xs, a1);
          a2 += dg;
          a3 = fmaf(dg, xs, a3);
          xv[c][j] = xs;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    const float mean = a0 * inv_h;                                   // of the shifted row
    const float rstd = rsqrtf(fmaxf(a1 * inv_h - mean * mean, 0.f) + eps);
    const float s1 = a2 * inv_h;
    const float s2 = rstd * (a3 - mean * a2) * inv_h;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        float g[8], o[8];
        load8<float>(gamma + ch * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[c][j] - mean) * rstd;
          acc_g[c][j] = fmaf(dv[c][j], xh, acc_g[c][j]);
          acc_b[c][j] += dv[c][j];
          o[j] = rstd * (dv[c][j] * g[j] - s1 - xh * s2);
        }
        store8<T>(dx + row * H + ch * 8, o);
        if (out_drop) {
          apply_dropout8(o, drop, site, row, H, ch * 8);
          store8<T>(dx_drop + row * H + ch * 8, o);
        }
        if (dbias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc_bias[c][j] += o[j];
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) {
      store8<float>(my + ch * 8, acc_g[c]);
      store8<float>(my + H + ch * 8, acc_b[c]);
      store8<float>(my + 2 * H + ch * 8, acc_bias[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) {
    const int which = i / H, col = i - which * H;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[static_cast<size_t>(w) * 3 * H + i];
    atomicAdd(dst + col, s);
  }
}

// Embedding backward.  grid = (joint position s, batch part): every row a CTA touches has the SAME position id (positions
// depend on s only), so the position gradient is summed in registers / shared memory and leaves the CTA as one atomicAdd
// per column; token-type gradients (type varies per sample only between real text and padding) are summed with
// shared-memory atomics first.  The word-embedding rows are scattered with global atomics (ids are mostly distinct),
// image rows go to the projection gradient.  The first version did three global atomicAdds per element, 27 904 of them
// landing on each of the 2 x 768 token-type addresses: 0.41 ms at 158 GB/s.
template <typename T>
__global__ void __launch_bounds__(256) embed_bwd_scatter_kernel(const EmbedBwdArgs a, int bper) {
  extern __shared__ float sm_e[];          // [8][H] per-warp position partials | [TV][H] token-type sums
  const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = a.H, nch = H >> 3;
  float* s_pos = sm_e;
  float* s_type = sm_e + 8 * H;
  for (int i = threadIdx.x; i < a.TV * H; i += 256) s_type[i] = 0.f;
  __syncthreads();
  const bool is_img = s >= 1 && s <= a.N, is_txt = s >= a.A;
  auto clamped = [](long id, int hi) -> long { return (hi > 0 && (id < 0 || id >= hi)) ? 0 : id; };   // as embed_ln_fwd (which reports)
  int pos_id = 0;
  if (is_img) pos_id = static_cast<int>(clamped(a.region_idx[s - 1], a.P));
  else if (s == a.N + 1) pos_id = a.sep_pos;
  else if (is_txt) pos_id = s - a.A;
  float acc[kMaxChunks][8];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  const int b0 = blockIdx.y * bper, b1 = min(a.B, b0 + bper);
  for (int b = b0 + warp; b < b1; b += 8) {
    long word_id = -1;
    int type_id = a.prefix_type;
    T* proj_row = nullptr;
    if (s == 0) word_id = clamped(a.cls_tok[b], a.V);
    else if (is_img) proj_row = static_cast<T*>(a.d_proj) + (static_cast<long>(b) * a.N + (s - 1)) * H;
    else if (s == a.N + 1) word_id = clamped(a.sep_tok[b], a.V);
    else {
      word_id = clamped(a.input_ids[static_cast<long>(b) * a.T + (s - a.A)], a.V);
      type_id = static_cast<int>(clamped(a.segment[static_cast<long>(b) * a.T + (s - a.A)], a.TV));
    }
    const T* src = static_cast<const T*>(a.dsum) + (static_cast<long>(b) * a.L + s) * H;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        float v[8];
        load8<T>(src + ch * 8, v);
        if (proj_row) store8<T>(proj_row + ch * 8, v);
        else if (word_id != a.pad_id) atomic_add8(a.d_word + word_id * H + ch * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[c][j] += v[j];
          atomicAdd(s_type + type_id * H + ch * 8 + j, v[j]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) store8<float>(s_pos + warp * H + ch * 8, acc[c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_pos[w * H + i];
    atomicAdd(a.d_pos + static_cast<long>(pos_id) * H + i, t);
  }
  for (int i = threadIdx.x; i < a.TV * H; i += 256) {
    const float t = s_type[i];
    if (t != 0.f) atomicAdd(a.d_type + i, t);
  }
}

// out[c] += sum_r x[r, c]; grid = (col groups of 256, row splits); 8 warps per CTA stride over rows
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long ld, int rows, int cols, float* out) {
  __shared__ float red[8][256 + 8];
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < cols) {
    for (long r = static_cast<long>(blockIdx.y) * 8 + warp; r < rows; r += static_cast<long>(gridDim.y) * 8) {
      float v[8];
      load8<T>(x + r * ld + col, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < cols) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(out + blockIdx.x * 256 + c, s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                          const int64_t* __restrict__ idx, int n, int period,
                                                          long stride, int H) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const long srow = static_cast<long>(warp / period) * stride + idx[warp % period];
  const uint4* s = reinterpret_cast<const uint4*>(src + srow * H);
  uint4* d = reinterpret_cast<uint4*>(dst + static_cast<long>(warp) * H);
  const int nvec = H * static_cast<int>(sizeof(T)) / 16;
  for (int i = lane; i < nvec; i += 32) d[i] = s[i];
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                           const int64_t* __restrict__ idx, int n, int period,
                                                           long stride, int H, int add) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const long drow = static_cast<long>(warp / period) * stride + idx[warp % period];
  const T* s = src + static_cast<long>(warp) * H;
  T* d = dst + drow * H;
  for (int ch = lane; ch < (H >> 3); ch += 32) {
    float v[8];
    load8<T>(s + ch * 8, v);
    if (add) {
      float o[8];
      load8<T>(d + ch * 8, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += o[j];
    }
    store8<T>(d + ch * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dgelu_kernel(const T* __restrict__ dy, const T* __restrict__ pre, T* __restrict__ dx, long n8) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float a[8], p[8];
    load8<T>(dy + i * 8, a);
    load8<T>(pre + i * 8, p);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= gelu_erf_grad(p[j]);
    store8<T>(dx + i * 8, a);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long n8) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    load8<T>(a + i * 8, x);
    load8<T>(b + i * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= y[j];
    store8<T>(out + i * 8, x);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) tanh_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ out, T* __restrict__ dpre, long n, int add) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float y = to_f32<T>(out[i]);
    const float g = to_f32<T>(dout[i]) * (1.f - y * y);
    dpre[i] = from_f32<T>(add ? to_f32<T>(dpre[i]) + g : g);
  }
}

__global__ void cast_f2b_kernel(const float* __restrict__ s, bf16* __restrict__ d, long n) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void cast_b2f_kernel(const bf16* __restrict__ s, float* __restrict__ d, long n) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    d[i] = __bfloat162float(s[i]);
}

__global__ void mask_dump_kernel(const unsigned char* mode, const int* t_len, int B, int A, int L, unsigned char* out) {
  const long total = static_cast<long>(B) * L * L;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % L), q = static_cast<int>((i / L) % L), b = static_cast<int>(i / (static_cast<long>(L) * L));
    out[i] = mask_allowed(mode[b], q, k, A, t_len[b]) ? 1 : 0;
  }
}

// One CTA per sample: probe three cells to pick the mode, count row 0 for t_len, then verify every cell.
__global__ void __launch_bounds__(256) mask_classify_kernel(const int64_t* mask, int dims, int B, int A, int L,
                                                            unsigned char* mode_out, int* tlen_out, int* mismatches) {
  const int b = blockIdx.x;
  __shared__ int s_mode, s_tlen, s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (dims == 2) {  // [B, L] key-padding mask: bidirectional
    const int64_t* m = mask + static_cast<long>(b) * L;
    int c = 0;
    for (int k = threadIdx.x; k < L; k += blockDim.x) c += m[k] != 0;
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) { s_mode = MODE_BIDIR; s_tlen = s_cnt - A; s_cnt = 0; }
    __syncthreads();
    int bad = 0;
    for (int k = threadIdx.x; k < L; k += blockDim.x) bad += (m[k] != 0) != mask_allowed(MODE_BIDIR, 0, k, A, s_tlen);
    if (bad) atomicAdd(mismatches, bad);
  } else {
    const int64_t* m = mask + static_cast<long>(b) * L * L;
    __shared__ int s_cnt_first, s_cnt_diag;
    if (threadIdx.x == 0) { s_cnt_first = 0; s_cnt_diag = 0; }
    __syncthreads();
    int c = 0, c0 = 0, cd = 0;
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
      c += m[static_cast<long>(L - 1) * L + k] != 0;   // last text row
      c0 += m[static_cast<long>(A) * L + k] != 0;      // first text row
      if (k >= A) cd += m[static_cast<long>(k) * L + k] != 0;   // text rows that see themselves
    }
    atomicAdd(&s_cnt, c);
    atomicAdd(&s_cnt_first, c0);
    atomicAdd(&s_cnt_diag, cd);
    __syncthreads();
    if (threadIdx.x == 0) {
      const bool txt_sees_img = m[static_cast<long>(A) * L + 0] != 0;
      const bool img_sees_txt = m[0 * L + A] != 0;
      // Bidirectional rows are all identical; the auto-regressive block makes the first and last text rows differ
      const bool rows_identical = s_cnt == s_cnt_first;
      int md;
      if (!txt_sees_img) md = MODE_NONCROSS;
      else if (!img_sees_txt) md = MODE_S2S;
      else if (!rows_identical) md = MODE_BAR;
      else md = MODE_BIDIR;
      int tl = md == MODE_BIDIR ? s_cnt - A : L - A;
      // fine-tune variants (data_loader.py:394-408): padded text rows see the prefix only and not themselves, so the
      // diagonal counts the real text rows; without pads they coincide with the pre-training masks
      if ((md == MODE_S2S || md == MODE_BAR) && s_cnt_diag < L - A && s_cnt == A) {
        md = md == MODE_S2S ? MODE_S2S_FT : MODE_BAR_FT;
        tl = s_cnt_diag;
      }
      s_mode = md;
      s_tlen = tl;
    }
    __syncthreads();
    const int md = s_mode, tl = s_tlen;
    int bad = 0;
    for (long i = threadIdx.x; i < static_cast<long>(L) * L; i += blockDim.x)
      bad += (m[i] != 0) != mask_allowed(md, static_cast<int>(i / L), static_cast<int>(i % L), A, tl);
    if (bad) atomicAdd(mismatches, bad);
  }
  if (threadIdx.x == 0) { mode_out[b] = static_cast<unsigned char>(s_mode); tlen_out[b] = s_tlen; }
}

inline int rows_grid(int rows) { return (rows * 32 + 255) / 256; }

}  // namespace

#define MV_DISPATCH_T(f32, ...)                                  \
  do {                                                           \
    if (f32) { using T = float; __VA_ARGS__; } else { using T = bf16; __VA_ARGS__; } \
  } while (0)

static int check_h(int H) {
  MV_REQUIRE(H % 8 == 0 && H <= 32 * kMaxChunks * 8 && H > 0, "row kernels need H %% 8 == 0 and H <= 1024 (got %d)", H);
  return 0;
}

int embed_ln_fwd(const EmbedArgs& a, int f32, cudaStream_t s) {
  if (check_h(a.H)) return -1;
  const int rows = a.B * a.L;
  MV_DISPATCH_T(f32, (embed_ln_fwd_kernel<T><<<rows_grid(rows), 256, 0, s>>>(a)));
  MV_LAUNCH_CHECK();
  return 0;
}

int embed_bwd_scatter(const EmbedBwdArgs& a, int f32, cudaStream_t s) {
  if (check_h(a.H)) return -1;
  MV_REQUIRE(a.TV >= 1 && a.TV <= 8, "embed_bwd_scatter: type vocabulary %d not in 1..8", a.TV);
  int parts = a.B / 16;                       // ~16 samples (two rows per warp) per CTA
  if (parts < 1) parts = 1;
  if (parts > 8) parts = 8;
  const int bper = (a.B + parts - 1) / parts;
  const size_t smem = static_cast<size_t>(8 + a.TV) * a.H * sizeof(float);
  static bool attr = false;
  if (!attr) {
    MV_CUDA_CHECK(cudaFuncSetAttribute(embed_bwd_scatter_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
    MV_CUDA_CHECK(cudaFuncSetAttribute(embed_bwd_scatter_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
    attr = true;
  }
  dim3 grid(a.L, (a.B + bper - 1) / bper);
  MV_DISPATCH_T(f32, (embed_bwd_scatter_kernel<T><<<grid, 256, smem, s>>>(a, bper)));
  MV_LAUNCH_CHECK();
  return 0;
}

int ln_fwd(const void* x, void* y, const float* gamma, const float* beta, int rows, int H, float eps, int drop_on,
           uint32_t drop_site, const DropoutCfg& drop, int f32, cudaStream_t s, int x_f32, float* y32) {
  if (check_h(H)) return -1;
  if (rows <= 0) return 0;
  const dim3 grid(rows_grid(rows)), block(256);
  if (f32) {
    MV_CUDA_CHECK(launch_pdl(ln_fwd_kernel<float, float>, grid, block, 0, s, static_cast<const float*>(x), static_cast<float*>(y), y32, gamma,
                             beta, rows, H, eps, drop_on, drop_site, drop));
  } else if (x_f32) {
    MV_CUDA_CHECK(launch_pdl(ln_fwd_kernel<float, bf16>, grid, block, 0, s, static_cast<const float*>(x), static_cast<bf16*>(y), y32, gamma,
                             beta, rows, H, eps, drop_on, drop_site, drop));
  } else {
    MV_CUDA_CHECK(launch_pdl(ln_fwd_kernel<bf16, bf16>, grid, block, 0, s, static_cast<const bf16*>(x), static_cast<bf16*>(y), y32, gamma,
                             beta, rows, H, eps, drop_on, drop_site, drop));
  }
  MV_LAUNCH_CHECK();
  return 0;
}

template <typename TX, typename T, int NC>
static int ln_bwd_launch(const void* dy, const void* x, const float* gamma, void* dx, void* dx_drop, float* dgamma, float* dbeta,
                         float* dbias, int rows, int H, float eps, int in_drop, int out_drop, uint32_t drop_site,
                         const DropoutCfg& drop, const LnAltDrop& alt, int grid, size_t smem, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    MV_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<TX, T, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4));
    attr = true;
  }
  MV_CUDA_CHECK(launch_pdl(ln_bwd_kernel<TX, T, NC>, dim3(grid), dim3(256), smem, s, static_cast<const T*>(dy), static_cast<const TX*>(x),
                           gamma, static_cast<T*>(dx), static_cast<T*>(dx_drop), dgamma, dbeta, dbias, rows, H, eps, in_drop, out_drop,
                           drop_site, drop, alt));
  MV_LAUNCH_CHECK();
  return 0;
}

int ln_bwd(const void* dy, const void* x, const float* gamma, void* dx, void* dx_drop, float* dgamma, float* dbeta,
           float* dbias, int rows, int H, float eps, int in_drop, int out_drop, uint32_t drop_site,
           const DropoutCfg& drop, int f32, cudaStream_t s, int x_f32, const LnAltDrop* alt_in) {
  if (check_h(H)) return -1;
  if (rows <= 0) return 0;
  MV_REQUIRE(!out_drop || dx_drop != nullptr, "ln_bwd: out_drop needs dx_drop");
  LnAltDrop alt;
  if (alt_in) alt = *alt_in; else { alt.period = 0; alt.lo = 0; alt.hi = 0; alt.drop = drop; }
  int grid = (rows + 7) / 8;
  const int cap = device_sm_count();          // one 8-warp CTA per SM: the row + the prefetched row + 72 column sums in registers
  if (grid > cap) grid = cap;
  const size_t smem = static_cast<size_t>(8) * 3 * H * sizeof(float);
  // H <= 768 (BERT-base): three 8-element chunks per lane, which keeps the whole row + the prefetched next row in registers
#define MV_LN_BWD(TX, T, NC) \
  return ln_bwd_launch<TX, T, NC>(dy, x, gamma, dx, dx_drop, dgamma, dbeta, dbias, rows, H, eps, in_drop, out_drop, drop_site, drop, alt, grid, smem, s)
  if (H <= 768) {
    if (f32) MV_LN_BWD(float, float, 3);
    if (x_f32) MV_LN_BWD(float, bf16, 3);
    MV_LN_BWD(bf16, bf16, 3);
  }
  if (f32) MV_LN_BWD(float, float, kMaxChunks);
  if (x_f32) MV_LN_BWD(float, bf16, kMaxChunks);
  MV_LN_BWD(bf16, bf16, kMaxChunks);
#undef MV_LN_BWD
}

int colsum_add(const void* x, long ld, int rows, int cols, float* out, int f32, cudaStream_t s) {
  MV_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "colsum: cols and ld must be multiples of 8");
  if (rows <= 0) return 0;
  int ysplit = (rows + 63) / 64;
  if (ysplit > 64) ysplit = 64;
  dim3 grid((cols + 255) / 256, ysplit);
  if (f32) MV_CUDA_CHECK(launch_pdl(colsum_kernel<float>, grid, dim3(256), 0, s, static_cast<const float*>(x), ld, rows, cols, out));
  else MV_CUDA_CHECK(launch_pdl(colsum_kernel<bf16>, grid, dim3(256), 0, s, static_cast<const bf16*>(x), ld, rows, cols, out));
  MV_LAUNCH_CHECK();
  return 0;
}

int gather_rows(const void* src, void* dst, const int64_t* idx, int n, int period, long stride, int H, int f32,
                cudaStream_t s) {
  if (n <= 0) return 0;
  MV_REQUIRE(H % 8 == 0, "gather_rows: H %% 8");
  MV_DISPATCH_T(f32, (gather_rows_kernel<T><<<rows_grid(n), 256, 0, s>>>(static_cast<const T*>(src), static_cast<T*>(dst), idx, n,
                                                                         period, stride, H)));
  MV_LAUNCH_CHECK();
  return 0;
}

int scatter_rows(const void* src, void* dst, const int64_t* idx, int n, int period, long stride, int H, int add,
                 int f32, cudaStream_t s) {
  if (n <= 0) return 0;
  MV_REQUIRE(H % 8 == 0, "scatter_rows: H %% 8");
  MV_DISPATCH_T(f32, (scatter_rows_kernel<T><<<rows_grid(n), 256, 0, s>>>(static_cast<const T*>(src), static_cast<T*>(dst), idx, n,
                                                                          period, stride, H, add)));
  MV_LAUNCH_CHECK();
  return 0;
}

int dgelu_mul(const void* dy, const void* pre, void* dx, long n, int f32, cudaStream_t s) {
  MV_REQUIRE(n % 8 == 0, "dgelu: n %% 8");
  if (n <= 0) return 0;
  const long n8 = n / 8;
  int grid = static_cast<int>((n8 + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  MV_DISPATCH_T(f32, (dgelu_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(dy), static_cast<const T*>(pre), static_cast<T*>(dx), n8)));
  MV_LAUNCH_CHECK();
  return 0;
}

int mul_elem(const void* a, const void* b, void* out, long n, int f32, cudaStream_t s) {
  MV_REQUIRE(n % 8 == 0, "mul_elem: n %% 8");
  if (n <= 0) return 0;
  const long n8 = n / 8;
  int grid = static_cast<int>((n8 + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  MV_DISPATCH_T(f32, (mul_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(a), static_cast<const T*>(b), static_cast<T*>(out), n8)));
  MV_LAUNCH_CHECK();
  return 0;
}

int tanh_bwd(const void* d_out, const void* out, void* d_pre, long n, int add, int f32, cudaStream_t s) {
  if (n <= 0) return 0;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  MV_DISPATCH_T(f32, (tanh_bwd_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(d_out), static_cast<const T*>(out), static_cast<T*>(d_pre), n, add)));
  MV_LAUNCH_CHECK();
  return 0;
}

int cast_f32_to_bf16(const float* src, bf16* dst, long n, cudaStream_t s) {
  if (n <= 0) return 0;
  cast_f2b_kernel<<<148 * 8, 256, 0, s>>>(src, dst, n);
  MV_LAUNCH_CHECK();
  return 0;
}
int cast_bf16_to_f32(const bf16* src, float* dst, long n, cudaStream_t s) {
  if (n <= 0) return 0;
  cast_b2f_kernel<<<148 * 8, 256, 0, s>>>(src, dst, n);
  MV_LAUNCH_CHECK();
  return 0;
}

int mask_dump(const unsigned char* mode, const int* t_len, int B, int A, int L, unsigned char* out, cudaStream_t s) {
  mask_dump_kernel<<<148 * 4, 256, 0, s>>>(mode, t_len, B, A, L, out);
  MV_LAUNCH_CHECK();
  return 0;
}

int mask_classify(const int64_t* mask, int dims, int B, int A, int L, unsigned char* mode, int* t_len,
                  int* mismatches, cudaStream_t s) {
  MV_REQUIRE(dims == 2 || dims == 3, "mask_classify: mask must be [B,L] or [B,L,L]");
  MV_CUDA_CHECK(cudaMemsetAsync(mismatches, 0, sizeof(int), s));
  mask_classify_kernel<<<B, 256, 0, s>>>(mask, dims, B, A, L, mode, t_len, mismatches);
  MV_LAUNCH_CHECK();
  return 0;
}

}  // namespace mv
