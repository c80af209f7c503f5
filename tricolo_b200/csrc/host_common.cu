#include "host_common.h"

#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

namespace tcl {

static thread_local char g_err[512] = "ok";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int require_sm100() {
  static std::mutex mu;
  static int cached[64];  // 0 unknown, 1 ok, 2 bad
  int dev = 0;
  TCL_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return set_error(TCL_ERR_BAD_ARCH, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(mu);
  if (cached[dev] == 0) {
    int major = 0, minor = 0;
    TCL_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    TCL_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    cached[dev] = (major == 10 && minor == 0) ? 1 : 2;
    if (cached[dev] == 2)
      return set_error(TCL_ERR_BAD_ARCH,
                       "tricolo_b200 is sm_100a only; device %d is sm_%d%d (no fallback path)", dev,
                       major, minor);
  }
  if (cached[dev] == 2)
    return set_error(TCL_ERR_BAD_ARCH, "tricolo_b200 is sm_100a only (device %d)", dev);
  return TCL_OK;
}

bool pdl_enabled() {
  const char* e = getenv("TRICOLO_B200_PDL");
  return !(e && e[0] == '0');
}

int ensure_dyn_smem_impl(const void* func, int bytes) {
  struct Entry { const void* func; int dev; int bytes; };
  static std::mutex mu;
  static Entry table[256];
  static int n_entries = 0;
  int dev = 0;
  TCL_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  Entry* hit = nullptr;
  for (int i = 0; i < n_entries; ++i)
    if (table[i].func == func && table[i].dev == dev) { hit = &table[i]; break; }
  if (hit && hit->bytes >= bytes) return TCL_OK;
  TCL_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (hit) hit->bytes = bytes;
  else if (n_entries < 256) table[n_entries++] = Entry{func, dev, bytes};  // table full: set it every time (still correct)
  return TCL_OK;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  });
  return fn;
}

int make_tmap_2d_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  TCL_REQUIRE(fn != nullptr, TCL_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  TCL_REQUIRE(aligned_to(base, 16), TCL_ERR_BAD_ALIGN, "TMA base pointer must be 16-byte aligned");
  TCL_REQUIRE((row_stride_elems * 2) % 16 == 0, TCL_ERR_BAD_ALIGN,
              "TMA row stride must be a multiple of 16 bytes (got %llu elements)",
              (unsigned long long)row_stride_elems);
  TCL_REQUIRE(box_cols * 2 == 128, TCL_ERR_BAD_ARG, "128-byte swizzle needs a 64-element inner box");
  TCL_REQUIRE(box_rows >= 1 && box_rows <= 256, TCL_ERR_BAD_ARG, "box rows out of range");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  // The 16-bit payload is moved as opaque bits; BFLOAT16 vs FLOAT16 only matters
  // for OOB NaN fill, which is not used.
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TCL_REQUIRE(r == CUDA_SUCCESS, TCL_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d",
              (int)r);
  return TCL_OK;
}

int make_tmap_3d_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t n_chunks, uint64_t row_stride_elems,
                       uint32_t box_rows, uint32_t box_chunks) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  TCL_REQUIRE(fn != nullptr, TCL_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  TCL_REQUIRE(aligned_to(base, 16) && (row_stride_elems * 2) % 16 == 0, TCL_ERR_BAD_ALIGN, "TMA 3D: 16-byte alignment");
  TCL_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_chunks >= 1 && box_chunks <= 256 && n_chunks >= 1, TCL_ERR_BAD_ARG,
              "TMA 3D: box out of range");
  cuuint64_t gdim[3] = {64, rows, n_chunks};
  cuuint64_t gstride[2] = {row_stride_elems * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_chunks};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TCL_REQUIRE(r == CUDA_SUCCESS, TCL_ERR_DRIVER, "cuTensorMapEncodeTiled (3D) failed with CUresult %d", (int)r);
  return TCL_OK;
}

int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                     uint32_t box_cols) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  TCL_REQUIRE(fn != nullptr, TCL_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  TCL_REQUIRE(aligned_to(base, 16) && (cols * 4) % 16 == 0, TCL_ERR_BAD_ALIGN, "TMA fp32 tensor: 16-byte alignment");
  TCL_REQUIRE(box_cols * 4 == 128, TCL_ERR_BAD_ARG, "128-byte swizzle needs a 32-element fp32 inner box");
  TCL_REQUIRE(box_rows >= 1 && box_rows <= 256, TCL_ERR_BAD_ARG, "box rows out of range");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TCL_REQUIRE(r == CUDA_SUCCESS, TCL_ERR_DRIVER, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return TCL_OK;
}

}  // namespace tcl
