// mask.cuh — MedViLL self-attention mask as a closed-form predicate of (mode, q, k, A, t_len).
// Restates the tensor construction at /root/reference/data/dataset_origin.py:138-176; the additive
// -10000 of models/cxrbert_origin.py:82-83 underflows to exactly 0 after softmax in fp32, so "masked" == "p = 0".
//   A      = num_image_embeds + 2   ([CLS] + regions + [SEP] prefix)
//   t_len  = real text length including its trailing [SEP] (only the Bidirectional mode and the fine-tune variants
//            depend on it)
// Fine-tune variants (report generation, Downstream_task/report_generation_and_vqa/sc/data_loader.py:394-408): the
// auto-regressive block ends at the REAL text end (second_end = A + t_len) and padded query rows see the prefix only.
#pragma once

namespace mv {

enum MaskMode : int { MODE_BIDIR = 0, MODE_S2S = 1, MODE_BAR = 2, MODE_NONCROSS = 3, MODE_S2S_FT = 4, MODE_BAR_FT = 5 };

__host__ __device__ __forceinline__ bool mask_allowed(int mode, int q, int k, int A, int t_len) {
  switch (mode) {
    case MODE_BIDIR: return k < A + t_len;                       // dataset_origin.py:138-139,169-176
    case MODE_S2S: return k < A || (q >= A && k <= q);           // :141-148
    case MODE_BAR: return q < A || k < A || k <= q;              // :158-161
    case MODE_S2S_FT: return k < A || (q >= A && q < A + t_len && k <= q);            // data_loader.py:405-408
    case MODE_BAR_FT: return q < A || k < A || (q < A + t_len && k <= q);             // data_loader.py:398-402
    default: return (q < A) == (k < A);                          // :163-167 (Non-cross)
  }
}

// Every mode's allowed key set for a query row is ONE interval [lo, hi): the kernels test (unsigned)(k - lo) < hi - lo.
//   Bidirectional [0, A + t_len) | Seq2Seq q<A: [0, A), else [0, q] | BAR q<A: [0, L), else [0, q] | Non-cross [0, A) or [A, L)
//   fine-tune Seq2Seq / BAR: as above for real text rows, [0, A) for padded rows (q >= A + t_len)
__host__ __device__ __forceinline__ void mask_row_interval(int mode, int q, int A, int t_len, int L, int& lo, int& hi) {
  lo = 0;
  switch (mode) {
    case MODE_BIDIR: hi = A + t_len; break;
    case MODE_S2S: hi = q < A ? A : q + 1; break;
    case MODE_BAR: hi = q < A ? L : q + 1; break;
    case MODE_S2S_FT: hi = (q < A || q >= A + t_len) ? A : q + 1; break;
    case MODE_BAR_FT: hi = q < A ? L : (q >= A + t_len ? A : q + 1); break;
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
    // real text rows of the tile are [max(q_lo, A), min(q_hi, A + t_len - 1)]: the largest one sees keys up to itself
    case MODE_S2S_FT: return k_lo < A || (q_lo < A + t_len && q_hi >= A && k_lo <= (q_hi < A + t_len - 1 ? q_hi : A + t_len - 1));
    case MODE_BAR_FT: return q_lo < A || k_lo < A || (q_lo < A + t_len && k_lo <= (q_hi < A + t_len - 1 ? q_hi : A + t_len - 1));
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
    case MODE_S2S_FT: return k_hi < A || (q_lo >= A && q_hi < A + t_len && k_hi <= q_lo);
    case MODE_BAR_FT: return q_hi < A || k_hi < A || (q_lo >= A && q_hi < A + t_len && k_hi <= q_lo);
    default: return (q_hi < A && k_hi < A) || (q_lo >= A && k_lo >= A);
  }
}

}  // namespace mv
