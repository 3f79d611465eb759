// Error reporting and device queries for the tcow_b200 C ABI (include/tcow_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include <cuda.h>

#include "tcow_internal.h"

namespace tcow {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (c == 0) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c > 0 ? c : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Resolved through the runtime so that the library has no link-time dependency on libcuda (it must load on
// a machine without a driver for the symbol-export test).
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rank-D tensor, dims[0] innermost (contiguous); strides_bytes[i] = byte stride of dim i+1; 128-byte swizzle.
int make_tmap_nd(void* map, bool is_f32, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TCOW_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return set_error(TCOW_ERR_ARG, "tensor map: base must be 16-byte aligned");
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] & 15) return set_error(TCOW_ERR_ARG, "tensor map: strides must be multiples of 16 bytes");
    st[i] = strides_bytes[i];
  }
  CUresult r = fn(static_cast<CUtensorMap*>(map), is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  rank, const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TCOW_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

// 2-D row-major tensor [rows, inner] with a row pitch; box = [box_rows, box_inner].
int make_tmap_2d(void* map, bool is_f32, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_elems,
                 uint32_t box_inner, uint32_t box_rows) {
  const uint64_t dims[2] = {inner, rows};
  const uint64_t strides[1] = {pitch_elems * (is_f32 ? 4u : 2u)};
  const uint32_t box[2] = {box_inner, box_rows};
  return make_tmap_nd(map, is_f32, ptr, 2, dims, strides, box);
}

// 4-D view of the patch rows of a token-row matrix: (column, t, n, b) -> ((b*N+n)*T+t)*ld + column; box = 64 columns x
// box_n tokens of one frame (rows T apart), SWIZZLE_128B — what the spatial-attention kernels load and store through.
int make_patch_tmap(void* m, const void* qkv, int64_t ld, int cols, int B, int N, int T, int box_n) {
  const uint64_t dims[4] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(T), static_cast<uint64_t>(N),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(T) * ld * 2,
                               static_cast<uint64_t>(N) * T * ld * 2};
  const uint32_t box[4] = {64, 1, static_cast<uint32_t>(box_n), 1};
  return make_tmap_nd(m, false, qkv, 4, dims, strides, box);
}

}  // namespace tcow

extern "C" int tcow_abi_version(void) { return 1; }
extern "C" const char* tcow_last_error(void) { return tcow::g_err; }
extern "C" int tcow_check_device(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return tcow::set_error(TCOW_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
  if (major != 10)
    return tcow::set_error(TCOW_ERR_ARCH, "tcow_b200 requires a Blackwell sm_100 device (found sm_%d0); no fallback", major);
  return 0;
}
