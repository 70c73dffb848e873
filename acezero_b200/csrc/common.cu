// Host-side plumbing shared by all translation units: error strings, tensor-map encoding, device queries.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>


namespace acez {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? ACEZ_ERR_NO_DEVICE : ACEZ_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return ACEZ_ERR_NO_DEVICE;
  }
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5];
  cuuint32_t e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u, stride0 %llu)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return ACEZ_ERR_CUDA;
  }
  return ACEZ_OK;
}

int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}

}  // namespace acez

extern "C" {

const char* acez_last_error(void) { return acez::g_err; }

int acez_version(void) { return ACEZ_VERSION; }

int acez_device_check(void) {
  static bool ok = false;  // cached: the check must not issue device queries inside a CUDA-graph capture
  if (ok) return ACEZ_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    acez::set_error("no CUDA device visible (%s)", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    return ACEZ_ERR_NO_DEVICE;
  }
  int dev = 0, major = 0, minor = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    acez::set_error("device compute capability %d.%d is not sm_100a; this library has no other code path", major,
                    minor);
    return ACEZ_ERR_UNSUPPORTED;
  }
  ok = true;
  return ACEZ_OK;
}

}  // extern "C"
