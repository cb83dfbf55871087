#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace orvb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return ORVB_ECUDA;
  ORVB_REQUIRE(box_cols * 2 == 128, ORVB_ESHAPE, "tensor map: box inner extent must be 128 bytes");
  ORVB_REQUIRE(box_rows >= 1 && box_rows <= 256, ORVB_ESHAPE, "tensor map: box rows %u out of range", box_rows);
  ORVB_REQUIRE((ld * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(base) % 16 == 0, ORVB_ESHAPE,
               "tensor map: base and row pitch must be 16-byte aligned");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ORVB_REQUIRE(r == CUDA_SUCCESS, ORVB_ECUDA, "cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu) failed: %d",
               (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, (int)r);
  return ORVB_OK;
}

// Unswizzled [box_rows x box_cols] boxes (rows packed at box_cols * 2 bytes in shared memory): the narrow last column
// unit of a GEMM tile whose width is not a multiple of 64.
int make_tmap_2d_bf16_plain(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                            uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return ORVB_ECUDA;
  ORVB_REQUIRE(box_cols >= 8 && box_cols % 8 == 0 && box_cols <= 128, ORVB_ESHAPE,
               "tensor map: plain box needs a multiple of 8 columns (got %u)", box_cols);
  ORVB_REQUIRE(box_rows >= 1 && box_rows <= 256, ORVB_ESHAPE, "tensor map: box rows %u out of range", box_rows);
  ORVB_REQUIRE((ld * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(base) % 16 == 0, ORVB_ESHAPE,
               "tensor map: base and row pitch must be 16-byte aligned");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ORVB_REQUIRE(r == CUDA_SUCCESS, ORVB_ECUDA, "cuTensorMapEncodeTiled(2d plain rows=%llu cols=%llu ld=%llu) failed: %d",
               (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, (int)r);
  return ORVB_OK;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols,
                      uint64_t ld_row, uint64_t ld_batch, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return ORVB_ECUDA;
  ORVB_REQUIRE(box_cols * 2 == 128, ORVB_ESHAPE, "tensor map: box inner extent must be 128 bytes");
  ORVB_REQUIRE(box_rows >= 1 && box_rows <= 256, ORVB_ESHAPE, "tensor map: box rows %u out of range", box_rows);
  ORVB_REQUIRE((ld_row * 2) % 16 == 0 && (ld_batch * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(base) % 16 == 0,
               ORVB_ESHAPE, "tensor map: base and pitches must be 16-byte aligned");
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {ld_row * 2, ld_batch * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ORVB_REQUIRE(r == CUDA_SUCCESS, ORVB_ECUDA, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return ORVB_OK;
}

int make_tmap_4d_bf16(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t t, uint32_t box_w,
                      uint32_t box_h) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return ORVB_ECUDA;
  ORVB_REQUIRE(c % 8 == 0 && reinterpret_cast<uintptr_t>(base) % 16 == 0, ORVB_ESHAPE,
               "tensor map: channels must be a multiple of 8 and the base 16-byte aligned");
  ORVB_REQUIRE(box_w >= 1 && box_h >= 1 && box_w * box_h <= 256, ORVB_ESHAPE, "tensor map: bad 4-D box");
  cuuint64_t gdim[4] = {c, w, h, t};
  cuuint64_t gstride[3] = {c * 2, w * c * 2, h * w * c * 2};
  cuuint32_t box[4] = {64, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ORVB_REQUIRE(r == CUDA_SUCCESS, ORVB_ECUDA, "cuTensorMapEncodeTiled(4d c=%llu w=%llu h=%llu t=%llu) failed: %d",
               (unsigned long long)c, (unsigned long long)w, (unsigned long long)h, (unsigned long long)t, (int)r);
  return ORVB_OK;
}

static int g_cc_major = -1;
static int g_sms = 0;

static int query_device() {
  if (g_cc_major >= 0) return ORVB_OK;
  int dev = 0;
  ORVB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ORVB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  g_cc_major = prop.major;
  g_sms = prop.multiProcessorCount;
  return ORVB_OK;
}

int check_arch() {
  int rc = query_device();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(g_cc_major == 10, ORVB_EARCH,
               "liborv_b200 needs a compute-capability 10.x device (B200, sm_100a); found major=%d", g_cc_major);
  return ORVB_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ORVB_PDL");
    v = (e != nullptr) ? atoi(e) : 1;
  }
  return v != 0;
}

int sm_count() {
  if (query_device() != ORVB_OK) return 148;
  return g_sms > 0 ? g_sms : 148;
}

}  // namespace orvb

extern "C" int orvb_version(void) { return ORVB_VERSION; }
extern "C" const char* orvb_last_error(void) { return orvb::get_error(); }
extern "C" int orvb_check_device(void) { return orvb::check_arch(); }
