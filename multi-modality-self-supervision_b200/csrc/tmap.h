// tmap.h — host-side TMA tensor-map encoding without linking libcuda (the driver entry point is
// resolved through the runtime, so the library still dlopen()s on a GPU-less build box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mv {

enum TmapDtype { TMAP_BF16 = 0, TMAP_F32 = 1 };

// 2-D row-major tensor: `inner` contiguous elements per row, `outer` rows, rows `row_stride_bytes` apart.
// Box = box_inner x box_outer elements, SWIZZLE_128B (box_inner * elem_size must be <= 128 B), OOB reads -> 0.
int tmap_encode_2d(CUtensorMap* out, TmapDtype dt, const void* base, uint64_t inner, uint64_t outer,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

// 3-D variant (attention: [batch*rows, heads.., cols] style views); strides in bytes for dims 1 and 2.
int tmap_encode_3d(CUtensorMap* out, TmapDtype dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                   uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);
// 4-D tiled map with caller-chosen strides (they may OVERLAP: the stem convolution reads 64-element windows that advance by
// 16 elements per output pixel); box = box0 x box1 x 1 x 1
int tmap_encode_4d(CUtensorMap* out, TmapDtype dt, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                   uint32_t box0, uint32_t box1);

}  // namespace mv
