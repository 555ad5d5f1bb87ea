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

}  // namespace mts
