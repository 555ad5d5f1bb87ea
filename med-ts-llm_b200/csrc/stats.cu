// stats.cu — the numbers of the "input statistics" prompt (models/medtsllm.py:441-495, calcute_lags :530-538) computed
// on the device in two launches and read back by the host in ONE packed copy (the reference issues five `.tolist()`
// device syncs per batch: min, max, median, trend, lags).
//
//   per (sample, feature) series x[0..T):  min, max, lower median (torch.median), trend = sign of sum(diff(x)),
//   circular autocorrelation corr[k] = sum_t x[t] x[(t+k) mod T]  (what irfft(rfft(x) conj(rfft(x))) evaluates for even T),
//   per sample: mean of corr over the features, indices of its 5 largest values.
//
// corr[k] == corr[T-k] exactly in exact arithmetic; the reference's FFT leaves the order inside such a pair to rounding
// noise.  Here corr is accumulated in fp64 for k <= T/2 and mirrored, ties resolve to the smaller index.
// Bytes: 4*T per series read once (the series lives in shared memory); the T^2/2 multiply-adds per series are fp64.
#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// grid (C_sel, B); x [B, T, C] fp32; feature f = f0 + blockIdx.x
//   stats [B, C_sel, 4] fp32 = (min, max, median, trend as 0/1);  corr [B, C_sel, T] fp64
__global__ void __launch_bounds__(256)
input_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, double* __restrict__ corr, int T, int C,
                   int f0, int C_sel) {
  extern __shared__ __align__(16) float series[];     // T floats
  __shared__ double red[32];
  __shared__ float s_min[8], s_max[8];
  const int b = blockIdx.y, fi = blockIdx.x, f = f0 + fi;
  const float* xb = x + (int64_t)b * T * C + f;
  float mn = INFINITY, mx = -INFINITY;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float v = xb[(int64_t)t * C];
    series[t] = v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
  __syncthreads();
  float* out = stats + ((int64_t)b * C_sel + fi) * 4;
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fminf(mn, s_min[w]); mx = fmaxf(mx, s_max[w]); }
    out[0] = mn;
    out[1] = mx;
  }
  // trend: sign of the sum of the fp32 first differences (xs.diff(dim=1).sum(dim=1) > 0), summed in fp64
  double ds = 0.0;
  for (int t = threadIdx.x; t + 1 < T; t += blockDim.x) ds += (double)(series[t + 1] - series[t]);
  ds = block_sum_d(ds, red);
  if (threadIdx.x == 0) out[3] = ds > 0.0 ? 1.0f : 0.0f;
  // lower median = order statistic (T-1)/2 (torch.median): the element with #less <= k < #less-or-equal
  const int kth = (T - 1) / 2;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const float v = series[i];
    int less = 0, leq = 0;
    for (int j = 0; j < T; ++j) {
      const float u = series[j];
      less += (u < v);
      leq += (u <= v);
    }
    if (less <= kth && kth < leq) out[2] = v;            // ties write the same value
  }
  // circular autocorrelation, k <= T/2 computed, the rest mirrored
  double* cb = corr + ((int64_t)b * C_sel + fi) * T;
  for (int k = threadIdx.x; k <= T / 2; k += blockDim.x) {
    double acc = 0.0;
    int j = k;
    for (int t = 0; t < T; ++t) {
      acc += (double)series[t] * (double)series[j];
      if (++j == T) j = 0;
    }
    cb[k] = acc;
    if (k > 0) cb[T - k] = acc;
  }
}

// grid (B); mean over the C_sel features of corr -> indices of the n_lags largest values (descending; ties -> smaller index)
__global__ void __launch_bounds__(256)
top_lags_kernel(const double* __restrict__ corr, int* __restrict__ lags, int T, int C_sel, int n_lags) {
  extern __shared__ __align__(16) double meanc[];      // T doubles
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < T; k += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < C_sel; ++c) s += corr[((int64_t)b * C_sel + c) * T + k];
    meanc[k] = s / C_sel;
  }
  __syncthreads();
  for (int r = 0; r < n_lags; ++r) {
    double best = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = threadIdx.x; k < T; k += blockDim.x) {
      const double v = meanc[k];
      if (v > best || (v == best && k < bi)) { best = v; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
      lags[(int64_t)b * n_lags + r] = bi;
      if (bi < T) meanc[bi] = -INFINITY;
    }
    __syncthreads();
  }
}

}  // namespace mts

using namespace mts;

extern "C" int mts_input_stats(const float* x, float* stats, double* corr, int* lags, int B, int T, int C, int f0,
                               int C_sel, int n_lags, mts_stream_t stream_) {
  if (!x || !stats || !corr || !lags || B <= 0 || T <= 1 || C <= 0 || f0 < 0 || C_sel <= 0 || f0 + C_sel > C ||
      n_lags <= 0 || n_lags > T)
    return set_error(MTS_ERR_INVALID_ARG, "mts_input_stats: bad arguments");
  if ((size_t)T * 8 > 96 * 1024)
    return set_error(MTS_ERR_UNSUPPORTED, "mts_input_stats: series longer than 12288 steps do not fit in shared memory");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(top_lags_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(input_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(input stats)", e);
    attr = true;
  }
  input_stats_kernel<<<dim3(C_sel, B), 256, (size_t)T * 4, stream>>>(x, stats, corr, T, C, f0, C_sel);
  count_launch();
  int rc = check_launch("input_stats_kernel");
  if (rc) return rc;
  top_lags_kernel<<<B, 256, (size_t)T * 8, stream>>>(corr, lags, T, C_sel, n_lags);
  count_launch();
  return check_launch("top_lags_kernel");
}
