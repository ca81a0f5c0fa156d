// tmap.cu — cuTensorMapEncodeTiled via cudaGetDriverEntryPoint; error plumbing storage.
#include "tmap.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace mv {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int tmap_encode_2d_uncached(CUtensorMap* out, TmapDtype dt, const void* base, uint64_t inner, uint64_t outer,
                            uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

// Descriptor cache: a training step encodes the same ~600 maps (same workspace pointers and shapes) every step, four per
// GEMM launch; cuTensorMapEncodeTiled costs ~1 us of host time each.  Keyed by every argument of the encode; thread-local
// (the engine launches from one thread), bounded.
namespace {
struct TmapKey {
  const void* base; uint64_t inner, outer, stride; uint32_t box_inner, box_outer, dt;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && stride == o.stride && box_inner == o.box_inner &&
           box_outer == o.box_outer && dt == o.dt;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.inner * 0xC2B2AE3D27D4EB4Full) ^ (k.outer << 17) ^ (k.stride << 33) ^ (static_cast<uint64_t>(k.box_inner) << 7) ^
         (static_cast<uint64_t>(k.box_outer) << 23) ^ k.dt;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
}  // namespace

int tmap_encode_2d(CUtensorMap* out, TmapDtype dt, const void* base, uint64_t inner, uint64_t outer,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{base, inner, outer, row_stride_bytes, box_inner, box_outer, static_cast<uint32_t>(dt)};
  auto hit = cache.find(key);
  if (hit != cache.end()) { *out = hit->second; return 0; }
  const int rc = tmap_encode_2d_uncached(out, dt, base, inner, outer, row_stride_bytes, box_inner, box_outer);
  if (rc == 0) {
    if (cache.size() > 16384) cache.clear();
    cache.emplace(key, *out);
  }
  return rc;
}

int tmap_encode_2d_uncached(CUtensorMap* out, TmapDtype dt, const void* base, uint64_t inner, uint64_t outer,
                            uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  MV_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const uint32_t esz = dt == TMAP_BF16 ? 2 : 4;
  MV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  MV_REQUIRE(row_stride_bytes % 16 == 0, "TMA row stride %llu not a multiple of 16 B",
             (unsigned long long)row_stride_bytes);
  MV_REQUIRE(box_inner * esz <= 128 && box_outer <= 256, "TMA box %ux%u too large", box_inner, box_outer);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt == TMAP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: CUresult %d (inner=%llu outer=%llu stride=%llu)",
             (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes);
  return 0;
}

int tmap_encode_3d(CUtensorMap* out, TmapDtype dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                   uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn enc = get_encode();
  MV_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const uint32_t esz = dt == TMAP_BF16 ? 2 : 4;
  MV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  MV_REQUIRE(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "TMA strides not multiples of 16 B");
  MV_REQUIRE(box0 * esz <= 128 && box1 <= 256 && box2 <= 256, "TMA box too large");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, dt == TMAP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r);
  return 0;
}

int tmap_encode_4d(CUtensorMap* out, TmapDtype dt, const void* base, const uint64_t dims_in[4], const uint64_t strides_bytes[3],
                   uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  MV_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const uint32_t esz = dt == TMAP_BF16 ? 2 : 4;
  MV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  MV_REQUIRE(strides_bytes[0] % 16 == 0 && strides_bytes[1] % 16 == 0 && strides_bytes[2] % 16 == 0, "TMA strides not multiples of 16 B");
  MV_REQUIRE(box0 * esz <= 128 && box1 <= 256, "TMA box too large");
  cuuint64_t dims[4] = {dims_in[0], dims_in[1], dims_in[2], dims_in[3]};
  cuuint64_t strides[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t box[4] = {box0, box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, dt == TMAP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed: CUresult %d (strides %llu %llu %llu)", (int)r,
             (unsigned long long)strides[0], (unsigned long long)strides[1], (unsigned long long)strides[2]);
  return 0;
}

static std::atomic<long> g_launches{0};
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MV_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// SMs the persistent GEMMs leave free (for NCCL's all-reduce CTAs while gradient buckets are in flight, engine.cu)
static int g_reserved_sms = 0;
void set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }
int gemm_sm_budget() {
  int n = device_sm_count() - g_reserved_sms;
  if (n < 2) n = 2;
  return n & ~1;            // CTA pairs
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace mv
