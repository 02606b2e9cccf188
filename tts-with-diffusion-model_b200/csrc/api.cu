// Error plumbing, device queries and TMA tensor-map encoding shared by all entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace vb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return VB200_ERR_CUDA;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p) {
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_2d(CUtensorMap* out, vb200_dtype dtype, const void* gptr, uint64_t inner,
                 uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return VB200_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) || (row_stride_bytes & 15)) {
    set_error("TMA operand must be 16-byte aligned (ptr=%p, row stride=%llu B)", gptr,
              static_cast<unsigned long long>(row_stride_bytes));
    return VB200_ERR_INVALID;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t elem[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == VB200_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == VB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                      : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, 2, const_cast<void*>(gptr), dims, strides,
                  box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu box=%ux%u)",
              static_cast<int>(r), static_cast<unsigned long long>(inner),
              static_cast<unsigned long long>(outer),
              static_cast<unsigned long long>(row_stride_bytes), box_inner, box_outer);
    return VB200_ERR_CUDA;
  }
  return VB200_OK;
}

struct TmapKey {
  const void* ptr; uint64_t inner, outer, stride; uint32_t box_inner, box_outer; int dtype;
  bool operator<(const TmapKey& o) const {
    return std::tie(ptr, inner, outer, stride, box_inner, box_outer, dtype) <
           std::tie(o.ptr, o.inner, o.outer, o.stride, o.box_inner, o.box_outer, o.dtype);
  }
};

int cached_tmap(CUtensorMap* out, vb200_dtype dtype, const void* ptr, uint64_t inner,
                uint64_t outer, uint64_t stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::map<TmapKey, CUtensorMap> cache;
  const TmapKey key{ptr, inner, outer, stride_bytes, box_inner, box_outer, static_cast<int>(dtype)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return VB200_OK; }
  }
  const int rc = make_tmap_2d(out, dtype, ptr, inner, outer, stride_bytes, box_inner, box_outer);
  if (rc != VB200_OK) return rc;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return VB200_OK;
}

}  // namespace vb200

extern "C" const char* vb200_last_error(void) { return vb200::g_err; }
extern "C" int vb200_version(void) { return 100; }
extern "C" int vb200_device_sms(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VB200_ERR_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return VB200_ERR_CUDA;
  return sms;
}
