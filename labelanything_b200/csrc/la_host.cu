// labelanything_b200 — host-side helpers shared by the C-ABI entry points:
// thread-local last-error string, run-time resolved cuTensorMapEncodeTiled, device queries.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "la_common.cuh"
#include "../../include/labelanything_b200.h"

namespace la {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static CUtensorMapSwizzle to_cu(Swizzle s) {
  switch (s) {
    case Swizzle::B32: return CU_TENSOR_MAP_SWIZZLE_32B;
    case Swizzle::B64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case Swizzle::B128: return CU_TENSOR_MAP_SWIZZLE_128B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int make_tensor_map_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, uint64_t inner, uint64_t outer,
                       uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, Swizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return LA_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  to_cu(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(2d) failed with CUresult %d (base=%p inner=%llu outer=%llu stride=%llu "
                   "box=%ux%u)",
                   (int)r, base, (unsigned long long)inner, (unsigned long long)outer,
                   (unsigned long long)outer_stride_bytes, box_inner, box_outer);
    return LA_ERR_CUDA;
  }
  return LA_OK;
}

int make_tensor_map_3d(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, uint64_t d0, uint64_t d1,
                       uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1,
                       uint32_t box2, Swizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return LA_ERR_CUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, dtype, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  to_cu(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
    return LA_ERR_CUDA;
  }
  return LA_OK;
}

int make_tensor_map_4d(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, const uint64_t dims[4],
                       const uint64_t strides_bytes[3], const uint32_t box[4], Swizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return LA_ERR_CUDA;
  }
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, dtype, 4, const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  to_cu(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r);
    return LA_ERR_CUDA;
  }
  return LA_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace la

extern "C" {

const char* la_last_error(void) { return la::g_last_error; }

int la_version(void) { return LA_B200_VERSION; }

int la_device_check(void) {
  int dev = 0;
  LA_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  LA_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  LA_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    la::set_last_error("labelanything_b200 kernels are built for sm_100a only; device %d is sm_%d%d", dev, major,
                       minor);
    return LA_ERR_UNSUPPORTED;
  }
  return LA_OK;
}

}  // extern "C"
