// Host-side helpers shared by every translation unit of liborv_b200.so: error convention,
// TMA tensor-map encoding (driver entry point resolved at run time so the library links
// without libcuda on the build box), and launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/orv_b200.h"

namespace orvb {

typedef __nv_bfloat16 bf16;

// thread-local last-error message, surfaced through orvb_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define ORVB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::orvb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ORVB_ECUDA;                                                                        \
    }                                                                                           \
  } while (0)

#define ORVB_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      ::orvb::set_error(__VA_ARGS__);   \
      return (code);                    \
    }                                   \
  } while (0)

// Encodes a 2-D bf16 row-major tensor [rows, cols] (cols contiguous, row pitch = ld elements)
// as a TMA descriptor with a {box_cols, box_rows} box and 128-byte swizzle.  box_cols must be 64
// (= 128 bytes of bf16, one swizzle span).  Out-of-bounds elements read as zero.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

// 3-D variant: [batch, rows, cols] with explicit element strides for rows and batch.
int make_tmap_2d_bf16_plain(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                            uint32_t box_rows, uint32_t box_cols);
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols,
                      uint64_t ld_row, uint64_t ld_batch, uint32_t box_rows, uint32_t box_cols);

// 4-D variant for channels-last activations [T, H, W, C] (C contiguous): dims (C, W, H, T), box {64, box_w, box_h, 1}.
int make_tmap_4d_bf16(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t t, uint32_t box_w,
                      uint32_t box_h);

// Kernel launch with optional programmatic dependent launch (see ptx.cuh pdl_wait): ORVB_PDL=0 disables it.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (pdl && pdl_enabled()) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

int check_arch();  // ORVB_OK only on compute capability 10.x
int sm_count();

}  // namespace orvb
