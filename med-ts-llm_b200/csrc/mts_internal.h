// mts_internal.h — host-side plumbing shared by the translation units of libmtsb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mts_b200.h"

namespace mts {

// Records a thread-local error message and returns `code`.
int set_error(int code, const char* fmt, ...);
int set_cuda_error(const char* what, cudaError_t e);
// cudaGetLastError() after a launch; MTS_OK or MTS_ERR_CUDA.
int check_launch(const char* kernel);
void count_launch();
// SM count of the current device (cached).
int num_sms();

// Builds (or fetches from the host-side cache) a 3-D bf16 tensor map
//   dims {k, rows, batch}, strides {ld, batch_stride} elements, box {box_k, box_rows, 1},
//   128-byte swizzle, zero fill out of bounds.
int get_tmap_bf16_3d(CUtensorMap* out, const void* base, int64_t k, int64_t rows, int64_t batch,
                     int64_t ld, int64_t batch_stride, int box_k, int box_rows);
// Same for elements of `elem_bytes` (2 = bf16, 4 = fp32 read by the tensor cores as TF32).
int get_tmap_3d(CUtensorMap* out, const void* base, int64_t k, int64_t rows, int64_t batch,
                int64_t ld, int64_t batch_stride, int box_k, int box_rows, int elem_bytes);

// Launches `kernel` with programmatic stream serialization (see pdl_wait() in ptx.cuh).  ONLY for kernels whose every
// thread executes pdl_wait() before touching global memory another kernel may have written.  mts_set_option("pdl", 0)
// turns the attribute off (plain stream order).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mts

// LAUNCH_PDL(kernel, grid, block, smem_bytes, stream, args...): launch_pdl + error return, for use inside the int-returning
// entry points / launchers.
#define LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                                         \
  do {                                                                                                             \
    cudaError_t pdl_e_ = mts::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream),  \
                                         __VA_ARGS__);                                                             \
    if (pdl_e_ != cudaSuccess) return mts::set_cuda_error("cudaLaunchKernelEx(" #kernel ")", pdl_e_);             \
  } while (0)
