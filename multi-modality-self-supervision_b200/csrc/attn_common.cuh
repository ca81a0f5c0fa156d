// attn_common.cuh — dropout indexing shared by the tcgen05 and SIMT attention kernels (forward and backward must
// regenerate the same keep-mask for attention-probability dropout, upstream BertSelfAttention.dropout).
#pragma once
#include "common.cuh"

namespace mv {

// Group index of the 16 consecutive keys [16*k16, 16*k16+16) of probability row (b, h, q).
__device__ __forceinline__ uint64_t attn_drop_group(int b, int nh, int h, int L, int q, int k16) {
  const uint64_t groups_per_row = static_cast<uint64_t>((L + 15) >> 4);
  return ((static_cast<uint64_t>(b) * nh + h) * L + q) * groups_per_row + k16;
}

__device__ __forceinline__ bool attn_keep(const DropoutCfg& d, uint32_t site, int b, int nh, int h, int L, int q, int k) {
  return keep16_bit(dropout_keep16(d, site, attn_drop_group(b, nh, h, L, q, k >> 4)), k & 15);
}

}  // namespace mv
