// mask.cuh — MedViLL self-attention mask as a closed-form predicate of (mode, q, k, A, t_len).
// Restates the tensor construction at /root/reference/data/dataset_origin.py:138-176; the additive
// -10000 of models/cxrbert_origin.py:82-83 underflows to exactly 0 after softmax in fp32, so "masked" == "p = 0".
//   A      = num_image_embeds + 2   ([CLS] + regions + [SEP] prefix)
//   t_len  = real text length including its trailing [SEP] (only the Bidirectional mode depends on it)
#pragma once

namespace mv {

enum MaskMode : int { MODE_BIDIR = 0, MODE_S2S = 1, MODE_BAR = 2, MODE_NONCROSS = 3 };

__host__ __device__ __forceinline__ bool mask_allowed(int mode, int q, int k, int A, int t_len) {
  switch (mode) {
    case MODE_BIDIR: return k < A + t_len;                       // dataset_origin.py:138-139,169-176
    case MODE_S2S: return k < A || (q >= A && k <= q);           // :141-148
    case MODE_BAR: return q < A || k < A || k <= q;              // :158-161
    default: return (q < A) == (k < A);                          // :163-167 (Non-cross)
  }
}

// Every mode's allowed key set for a query row is ONE interval [lo, hi): the kernels test (unsigned)(k - lo) < hi - lo.
//   Bidirectional [0, A + t_len) | Seq2Seq q<A: [0, A), else [0, q] | BAR q<A: [0, L), else [0, q] | Non-cross [0, A) or [A, L)
__host__ __device__ __forceinline__ void mask_row_interval(int mode, int q, int A, int t_len, int L, int& lo, int& hi) {
  lo = 0;
  switch (mode) {
    case MODE_BIDIR: hi = A + t_len; break;
    case MODE_S2S: hi = q < A ? A : q + 1; break;
    case MODE_BAR: hi = q < A ? L : q + 1; break;
    default: if (q < A) hi = A; else { lo = A; hi = L; } break;
  }
  if (hi > L) hi = L;
}

// true iff at least one (q, k) with q in [q_lo, q_hi], k in [k_lo, k_hi] (inclusive) is allowed: exact, so tiles for
// which this is false can be skipped without changing the result.
__host__ __device__ __forceinline__ bool tile_any_allowed(int mode, int q_lo, int q_hi, int k_lo, int k_hi, int A,
                                                          int t_len) {
  switch (mode) {
    case MODE_BIDIR: return k_lo < A + t_len;
    case MODE_S2S: return k_lo < A || (q_hi >= A && k_lo <= q_hi);
    case MODE_BAR: return q_lo < A || k_lo < A || k_lo <= q_hi;
    default: return (q_lo < A && k_lo < A) || (q_hi >= A && k_hi >= A);
  }
}

// true iff every pair in the tile is allowed (predicate evaluation can be skipped)
__host__ __device__ __forceinline__ bool tile_all_allowed(int mode, int q_lo, int q_hi, int k_lo, int k_hi, int A,
                                                          int t_len) {
  switch (mode) {
    case MODE_BIDIR: return k_hi < A + t_len;
    case MODE_S2S: return k_hi < A || (q_lo >= A && k_hi <= q_lo);
    case MODE_BAR: return q_hi < A || k_hi < A || k_hi <= q_lo;
    default: return (q_hi < A && k_hi < A) || (q_lo >= A && k_lo >= A);
  }
}

}  // namespace mv
