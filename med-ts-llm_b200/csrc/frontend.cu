// frontend.cu — K1+K2: RevIN statistics/normalisation, end-replication padding, patch unfolding and
// the 3-tap circular TokenEmbedding conv in ONE pass over the window (ref: models/layers/RevIN.py:
// 37-56; models/layers/embed.py:155-163, 186-197, 29-46; concat reshape models/medtsllm.py:276-279).
//
// HBM-bound by construction: a window [T,C] fp32 is read once with coalesced 16-byte loads into
// shared memory, every later access (statistics, normalisation, the 3*P taps) hits smem, and the
// patch embeddings are written once, fully coalesced.  Algorithmic bytes per window:
// 4*T*C (read) + 2*C*N*d_model (bf16 write) + 8*C (statistics).
#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

__device__ __forceinline__ float fe_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Patch index map (the bit-exact contract): padded sample index of element p of patch n.
__device__ __forceinline__ int patch_src_index(int n, int p, int S, int T) {
  const int t = n * S + p;
  return t < T ? t : T - 1;  // ReplicationPad1d((0, S)): repeat the last sample
}

// Loads x[b] ([T,C] contiguous) into smem and normalises it per channel.  Returns with xs = the
// normalised window and s_mean/s_std filled.  If mean_in != nullptr the statistics are taken from
// memory (backward), else computed (forward).
__device__ void load_and_normalise(const float* __restrict__ xb, float* xs, float* s_mean,
                                   float* s_std, int T, int C, float eps,
                                   const float* mean_in, const float* std_in) {
  const int n = T * C;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
    for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x)
      reinterpret_cast<float4*>(xs)[i] = __ldg(reinterpret_cast<const float4*>(xb) + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = __ldg(xb + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (mean_in) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) { s_mean[c] = mean_in[c]; s_std[c] = std_in[c]; }
  } else {
    for (int c = warp; c < C; c += nw) {
      float s = 0.0f;
      for (int t = lane; t < T; t += 32) s += xs[t * C + c];
      const float mean = fe_warp_sum(s) / T;
      float v = 0.0f;
      for (int t = lane; t < T; t += 32) { const float d = xs[t * C + c] - mean; v += d * d; }
      const float var = fe_warp_sum(v) / T;  // unbiased=False
      if (lane == 0) { s_mean[c] = mean; s_std[c] = sqrtf(var + eps); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % C;
    xs[i] = (xs[i] - s_mean[c]) / s_std[c];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
revin_patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ w_conv,
                         float* __restrict__ mean, float* __restrict__ stdev,
                         __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int T,
                         int C, int P, int S, int N, int dm, int concat, float eps) {
  extern __shared__ float fe_smem[];
  float* xs = fe_smem;                          // [T*C] (padded to a multiple of 4)
  float* ws = xs + ((T * C + 3) & ~3);          // [3][P][dm]  (tap, p, d): d fastest
  float* s_mean = ws + 3 * P * dm;              // [C]
  float* s_std = s_mean + C;                    // [C]
  const int b = blockIdx.x;

  for (int i = threadIdx.x; i < dm * P * 3; i += blockDim.x) {
    // w_conv is [dm][P][3]
    const int k = i % 3, pp = (i / 3) % P, d = i / (3 * P);
    ws[(k * P + pp) * dm + d] = __ldg(w_conv + i);
  }
  load_and_normalise(x + (int64_t)b * T * C, xs, s_mean, s_std, T, C, eps, nullptr, nullptr);
  if (blockIdx.y == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      mean[(int64_t)b * C + c] = s_mean[c];
      stdev[(int64_t)b * C + c] = s_std[c];
    }
  }
  // this CTA's slice of the patch axis
  const int per = (N + gridDim.y - 1) / gridDim.y;
  const int n_lo = blockIdx.y * per;
  const int n_hi = min(N, n_lo + per);
  const int total = (n_hi - n_lo) * C * dm;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int d = idx % dm;
    const int c = (idx / dm) % C;
    const int n = n_lo + idx / (dm * C);
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int nn = n + k - 1;  // circular over the patch axis
      nn = nn < 0 ? nn + N : (nn >= N ? nn - N : nn);
      const float* wk = ws + k * P * dm + d;
      for (int pp = 0; pp < P; ++pp)
        acc = fmaf(wk[pp * dm], xs[patch_src_index(nn, pp, S, T) * C + c], acc);
    }
    int64_t o;
    if (concat) o = ((int64_t)b * N + n) * ((int64_t)C * dm) + (int64_t)c * dm + d;
    else        o = (((int64_t)b * C + c) * N + n) * dm + d;
    if (out_bf16) out_bf16[o] = __float2bfloat16_rn(acc);
    if (out_f32) out_f32[o] = acc;
  }
}

__global__ void patch_gather_kernel(const float* __restrict__ x, float* __restrict__ patches, int B,
                                    int T, int C, int P, int S, int N) {
  const int64_t total = (int64_t)B * C * N * P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % P);
    const int n = (int)((i / P) % N);
    const int c = (int)((i / ((int64_t)P * N)) % C);
    const int b = (int)(i / ((int64_t)P * N * C));
    patches[i] = x[((int64_t)b * T + patch_src_index(n, p, S, T)) * C + c];
  }
}

// dW[d][p][k] = sum_{b,c,n} dout[b,n,c,d] * xn[b,c, src((n+k-1) mod N, p)]
__global__ void __launch_bounds__(256)
revin_patch_embed_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                             const float* __restrict__ stdev, const float* __restrict__ dout,
                             float* __restrict__ dw, int T, int C, int P, int S, int N, int dm,
                             int concat) {
  extern __shared__ float fe_smem[];
  float* xs = fe_smem;
  float* s_mean = xs + ((T * C + 3) & ~3);
  float* s_std = s_mean + C;
  const int b = blockIdx.x;
  load_and_normalise(x + (int64_t)b * T * C, xs, s_mean, s_std, T, C, 0.0f,
                     mean + (int64_t)b * C, stdev + (int64_t)b * C);
  const int entries = dm * P * 3;
  for (int e = threadIdx.x; e < entries; e += blockDim.x) {
    const int k = e % 3, pp = (e / 3) % P, d = e / (3 * P);
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) {
      for (int n = 0; n < N; ++n) {
        int nn = n + k - 1;
        nn = nn < 0 ? nn + N : (nn >= N ? nn - N : nn);
        int64_t o;
        if (concat) o = ((int64_t)b * N + n) * ((int64_t)C * dm) + (int64_t)c * dm + d;
        else        o = (((int64_t)b * C + c) * N + n) * dm + d;
        acc = fmaf(__ldg(dout + o), xs[patch_src_index(nn, pp, S, T) * C + c], acc);
      }
    }
    atomicAdd(dw + e, acc);
  }
}

__global__ void revin_denorm_kernel(float* __restrict__ y, const float* __restrict__ mean,
                                    const float* __restrict__ stdev, int64_t total, int T, int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t b = i / ((int64_t)T * C);
    y[i] = y[i] * stdev[b * C + c] + mean[b * C + c];
  }
}

static int fe_check(const char* name, int B, int T, int C, int P, int S) {
  if (B <= 0 || T <= 0 || C <= 0 || P <= 0 || S <= 0 || T + S < P)
    return set_error(MTS_ERR_INVALID_ARG, "%s: bad shape B=%d T=%d C=%d P=%d S=%d", name, B, T, C, P, S);
  return MTS_OK;
}

// ---------------------------------------------------------------------------------------------
// GPT4TS front end (ref: models/gpt4ts.py:126-138, :151-164, :200-212; models/layers/embed.py:29-46, 109-131):
// per-(sample, channel) normalisation over time, then per time step the 3-tap circular TokenEmbedding conv over
// TIME (in-channels = the C variables) + the fixed sinusoid table, written in the layout the next step wants:
//   mode 0  bf16, transposed [B, d_model, ld_t] — the K-major B operand of the time-axis Linear (forecasting)
//   mode 1  fp32 [B, T, D] residual-stream rows: embedding (+ zero pad to D) + wpe[t]   (segmentation tasks)
// grid (chunks, B): every CTA loads and normalises its sample's whole window (it is tiny) and emits a chunk of
// time steps.  HBM bytes per sample: 4*T*C read, 2*T*d_model (mode 0) or 4*T*D (mode 1) written.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gpt4ts_embed_kernel(const float* __restrict__ x, const float* __restrict__ w_conv, const float* __restrict__ pe,
                    const float* __restrict__ wpe, float* __restrict__ mean, float* __restrict__ stdev,
                    __nv_bfloat16* __restrict__ out_t, float* __restrict__ out_x, __nv_bfloat16* __restrict__ out_nt,
                    int T, int C, int d_model, int D, int ld_t, int mode, float eps) {
  extern __shared__ __align__(16) float fe_smem[];
  float* xs = fe_smem;
  float* s_mean = xs + ((T * C + 3) & ~3);
  float* s_std = s_mean + C;
  const int b = blockIdx.y;
  load_and_normalise(x + (int64_t)b * T * C, xs, s_mean, s_std, T, C, eps, nullptr, nullptr);
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) { mean[b * C + c] = s_mean[c]; stdev[b * C + c] = s_std[c]; }
  const int per = (T + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(T, t0 + per);
  const int width = mode == 0 ? d_model : D;
  for (int d = threadIdx.x; d < width; d += blockDim.x) {
    const float* wd = w_conv + (int64_t)d * C * 3;     // [d][c][k]
    for (int t = t0; t < t1; ++t) {
      float acc = 0.f;
      if (d < d_model) {
        const float* xp = xs + (t == 0 ? T - 1 : t - 1) * C;
        const float* xc = xs + t * C;
        const float* xn = xs + (t == T - 1 ? 0 : t + 1) * C;
        for (int c = 0; c < C; ++c) acc += wd[c * 3] * xp[c] + wd[c * 3 + 1] * xc[c] + wd[c * 3 + 2] * xn[c];
        acc += pe[(int64_t)t * d_model + d];
      }
      if (mode == 0) {
        out_t[((int64_t)b * d_model + d) * ld_t + t] = __float2bfloat16_rn(acc);
        if (out_nt) out_nt[((int64_t)b * T + t) * d_model + d] = __float2bfloat16_rn(acc);   // [B, T, d_model] (training)
      } else {
        out_x[((int64_t)b * T + t) * D + d] = acc + wpe[(int64_t)t * D + d];
      }
    }
  }
  if (mode == 0 && blockIdx.x == 0 && ld_t > T)   // zero the row padding (the GEMM's TMA extent stops at T anyway)
    for (int i = threadIdx.x; i < d_model * (ld_t - T); i += blockDim.x)
      out_t[((int64_t)b * d_model + i / (ld_t - T)) * ld_t + T + i % (ld_t - T)] = __float2bfloat16_rn(0.f);
}

// Weight gradient of that conv: dW[d, c, k] = sum_{b,t} denc(b, d, t) * xn[b, (t + k - 1) mod T, c], xn recomputed from
// the raw window and the saved statistics.  One CTA per output channel d, one thread per (c, k); deterministic.
__global__ void __launch_bounds__(128)
gpt4ts_conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdev,
                         const float* __restrict__ denc, float* __restrict__ dw, int B, int T, int C, int d_model,
                         int64_t sb, int64_t sd, int64_t st) {
  const int d = blockIdx.x;
  for (int j = threadIdx.x; j < 3 * C; j += blockDim.x) {
    const int c = j / 3, k = j - c * 3;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const float mu = mean[b * C + c], inv = 1.0f / stdev[b * C + c];
      const float* g = denc + b * sb + d * sd;          // element (b, d, t) at b*sb + d*sd + t*st
      const float* xb = x + (int64_t)b * T * C + c;
      for (int t = 0; t < T; ++t) {
        int ts = t + k - 1;
        ts = ts < 0 ? T - 1 : (ts >= T ? 0 : ts);
        acc += g[t * st] * ((xb[(int64_t)ts * C] - mu) * inv);
      }
    }
    dw[((int64_t)d * C + c) * 3 + k] = acc;
  }
}

}  // namespace mts

using namespace mts;

extern "C" int mts_gpt4ts_conv_wgrad(const float* x, const float* mean, const float* stdev, const float* denc,
                                     float* dw, int B, int T, int C, int d_model, int64_t sb, int64_t sd, int64_t st,
                                     mts_stream_t s) {
  if (!x || !mean || !stdev || !denc || !dw || B <= 0 || T <= 0 || C <= 0 || d_model <= 0 || sb <= 0 || sd <= 0 || st <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gpt4ts_conv_wgrad: bad args");
  gpt4ts_conv_wgrad_kernel<<<d_model, 128, 0, (cudaStream_t)s>>>(x, mean, stdev, denc, dw, B, T, C, d_model, sb, sd, st);
  count_launch();
  return check_launch("gpt4ts_conv_wgrad_kernel");
}

extern "C" int mts_gpt4ts_embed(const float* x, const float* w_conv, const float* pe, const float* wpe, float* mean,
                                float* stdev, uint16_t* out_t, float* out_x, uint16_t* out_nt, int B, int T, int C,
                                int d_model, int D, int ld_t, int mode, float eps, mts_stream_t s) {
  if (!x || !mean || !stdev || !w_conv || !pe || B <= 0 || T <= 0 || C <= 0 || d_model <= 0 || mode < 0 || mode > 1)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gpt4ts_embed: bad args");
  if (mode == 0 && (!out_t || ld_t < T))
    return set_error(MTS_ERR_INVALID_ARG, "mts_gpt4ts_embed: mode 0 needs out_t and ld_t >= T");
  if (mode == 1 && (!out_x || !wpe || D < d_model))
    return set_error(MTS_ERR_INVALID_ARG, "mts_gpt4ts_embed: mode 1 needs out_x, wpe and D >= d_model");
  const size_t smem = sizeof(float) * ((((size_t)T * C + 3) & ~(size_t)3) + 2 * (size_t)C);
  if (smem > 200 * 1024) return set_error(MTS_ERR_UNSUPPORTED, "mts_gpt4ts_embed: window too large");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gpt4ts_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gpt4ts embed)", e);
    smem_set = 200 * 1024;
  }
  int chunks = (2 * num_sms() + B - 1) / B;
  if (chunks > T) chunks = T;
  if (chunks < 1) chunks = 1;
  gpt4ts_embed_kernel<<<dim3(chunks, B), 256, smem, (cudaStream_t)s>>>(
      x, w_conv, pe, wpe, mean, stdev, reinterpret_cast<__nv_bfloat16*>(out_t), out_x,
      reinterpret_cast<__nv_bfloat16*>(out_nt), T, C, d_model, D, ld_t, mode, eps);
  count_launch();
  return check_launch("gpt4ts_embed_kernel");
}

extern "C" int mts_revin_patch_embed(const float* x, const float* w_conv, float* mean, float* stdev,
                                     uint16_t* out_bf16, float* out_f32, int B, int T, int C, int P,
                                     int S, int d_model, int concat_layout, float eps,
                                     mts_stream_t s) {
  int rc = fe_check("mts_revin_patch_embed", B, T, C, P, S);
  if (rc) return rc;
  if (!x || !w_conv || !mean || !stdev || (!out_bf16 && !out_f32) || d_model <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_revin_patch_embed: null pointer");
  const int N = (T + S - P) / S + 1;
  const size_t smem = sizeof(float) * (((size_t)T * C + 3 & ~(size_t)3) + 3 * (size_t)P * d_model + 2 * (size_t)C);
  if (smem > 200 * 1024)
    return set_error(MTS_ERR_UNSUPPORTED, "mts_revin_patch_embed: window of %d x %d floats exceeds shared memory", T, C);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(revin_patch_embed_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(frontend)", e);
    smem_set = 200 * 1024;
  }
  // split the patch axis so that B*splits CTAs cover the 148 SMs
  int splits = (num_sms() + B - 1) / B;
  if (splits > N) splits = N;
  if (splits < 1) splits = 1;
  dim3 grid(B, splits);
  revin_patch_embed_kernel<<<grid, 256, smem, (cudaStream_t)s>>>(
      x, w_conv, mean, stdev, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32, T, C, P, S, N,
      d_model, concat_layout, eps);
  count_launch();
  return check_launch("revin_patch_embed_kernel");
}

extern "C" int mts_patch_gather(const float* x, float* patches, int B, int T, int C, int P, int S,
                                mts_stream_t s) {
  int rc = fe_check("mts_patch_gather", B, T, C, P, S);
  if (rc) return rc;
  if (!x || !patches) return set_error(MTS_ERR_INVALID_ARG, "mts_patch_gather: null pointer");
  const int N = (T + S - P) / S + 1;
  const int64_t total = (int64_t)B * C * N * P;
  int64_t g = (total + 255) / 256;
  if (g > (int64_t)num_sms() * 16) g = (int64_t)num_sms() * 16;
  patch_gather_kernel<<<(int)g, 256, 0, (cudaStream_t)s>>>(x, patches, B, T, C, P, S, N);
  count_launch();
  return check_launch("patch_gather_kernel");
}

extern "C" int mts_revin_patch_embed_bwd(const float* x, const float* mean, const float* stdev,
                                         const float* dout, float* dw_conv, int B, int T, int C,
                                         int P, int S, int d_model, int concat_layout,
                                         mts_stream_t s) {
  int rc = fe_check("mts_revin_patch_embed_bwd", B, T, C, P, S);
  if (rc) return rc;
  if (!x || !mean || !stdev || !dout || !dw_conv || d_model <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_revin_patch_embed_bwd: null pointer");
  const int N = (T + S - P) / S + 1;
  const size_t smem = sizeof(float) * (((size_t)T * C + 3 & ~(size_t)3) + 2 * (size_t)C);
  if (smem > 200 * 1024)
    return set_error(MTS_ERR_UNSUPPORTED, "mts_revin_patch_embed_bwd: window too large");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(revin_patch_embed_bwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(frontend bwd)", e);
    smem_set = 200 * 1024;
  }
  cudaError_t e = cudaMemsetAsync(dw_conv, 0, sizeof(float) * 3 * (size_t)P * d_model, (cudaStream_t)s);
  if (e != cudaSuccess) return set_cuda_error("cudaMemsetAsync(dw_conv)", e);
  revin_patch_embed_bwd_kernel<<<B, 256, smem, (cudaStream_t)s>>>(x, mean, stdev, dout, dw_conv, T,
                                                                  C, P, S, N, d_model, concat_layout);
  count_launch();
  return check_launch("revin_patch_embed_bwd_kernel");
}

extern "C" int mts_revin_denorm(float* y, const float* mean, const float* stdev, int B, int T, int C,
                                mts_stream_t s) {
  if (!y || !mean || !stdev || B <= 0 || T <= 0 || C <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_revin_denorm: bad args");
  const int64_t total = (int64_t)B * T * C;
  int64_t g = (total + 255) / 256;
  if (g > (int64_t)num_sms() * 16) g = (int64_t)num_sms() * 16;
  revin_denorm_kernel<<<(int)g, 256, 0, (cudaStream_t)s>>>(y, mean, stdev, total, T, C);
  count_launch();
  return check_launch("revin_denorm_kernel");
}
