// rowops.cu — the HBM-bound row / element-wise kernels of the backbone: casts, transposes, RMSNorm,
// LayerNorm, row softmax, SwiGLU, prompt gather.  All are single-pass over HBM (the second
// pass of the reductions re-reads a row that is still L1/L2 resident), 16-byte vectorised, one
// warp-shuffle reduction tree per row.  Algorithmic bytes are stated per kernel.
#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for blockDim.x <= 1024; every thread receives the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.0f;
  return warp_sum(t);
}

// ------------------------------------------------------------------------------------------
// casts (bytes: 6 per element)
// ------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                     int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(in + i);
      const float4 b = *reinterpret_cast<const float4*>(in + i + 4);
      *reinterpret_cast<uint4*>(out + i) = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w),
                                                      pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    } else {
      for (int64_t j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
    }
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                     int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __bfloat162float(in[i]);
}

// 32x32 tiled transpose through padded smem; TIn in {float, bf16}, output bf16.
template <typename TIn>
__global__ void transpose_to_bf16_kernel(const TIn* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                         int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    float v = 0.0f;
    if (r < rows && c < cols) v = static_cast<float>(in[(int64_t)r * cols + c]);
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

// ------------------------------------------------------------------------------------------
// RMSNorm / LayerNorm: one CTA per row (bytes per row: 4*D read + 2*D (or 4*D) written)
// ------------------------------------------------------------------------------------------
template <bool kLayerNorm>
__global__ void __launch_bounds__(256, 6)
norm_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ y_bf16,
                 float* __restrict__ y_f32, int D, float eps) {
  pdl_wait();
  pdl_trigger();
  // One CTA per row.  The row (up to 4096 columns: 4 float4 per thread) is fetched ONCE, all loads in flight together,
  // and stays in registers through the reductions; wider rows re-read the remainder.  (The 16 KB of norm weights come
  // from L1/L2 in the last pass: caching them too costs the registers that keep 6 CTAs per SM resident.)
  constexpr int kCache = 4;
  __shared__ float red[32];
  const float* xr = x + (int64_t)blockIdx.x * ldx;
  const int D4 = D >> 2;  // D % 4 == 0 enforced on the host
  float4 v[kCache];
#pragma unroll
  for (int k = 0; k < kCache; ++k) {
    const int i = threadIdx.x + k * 256;
    v[k] = i < D4 ? reinterpret_cast<const float4*>(xr)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.0f, ss = 0.0f;
#pragma unroll
  for (int k = 0; k < kCache; ++k) {
    if (kLayerNorm) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  }
  for (int i = threadIdx.x + kCache * 256; i < D4; i += 256) {          // D > 4096 only
    const float4 t = reinterpret_cast<const float4*>(xr)[i];
    if (kLayerNorm) s += (t.x + t.y) + (t.z + t.w);
    ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  float mean = 0.0f, rstd;
  if (kLayerNorm) {
    mean = block_sum(s, red) / D;
    // second pass for the variance: matches torch's two-pass numerics better than E[x^2]-m^2
    float vs = 0.0f;
#pragma unroll
    for (int k = 0; k < kCache; ++k) {
      if (threadIdx.x + k * 256 < D4) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        vs += a * a + b * b + c * c + d * d;
      }
    }
    for (int i = threadIdx.x + kCache * 256; i < D4; i += 256) {
      const float4 t = reinterpret_cast<const float4*>(xr)[i];
      const float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
      vs += a * a + b * b + c * c + d * d;
    }
    rstd = rsqrtf(block_sum(vs, red) / D + eps);
  } else {
    rstd = rsqrtf(block_sum(ss, red) / D + eps);
  }
  auto emit = [&](int i, const float4& t, const float4& gw) {
    float4 o;
    o.x = (t.x - mean) * rstd * gw.x;
    o.y = (t.y - mean) * rstd * gw.y;
    o.z = (t.z - mean) * rstd * gw.z;
    o.w = (t.w - mean) * rstd * gw.w;
    if (kLayerNorm) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + i);
      o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
    }
    if (y_bf16)
      reinterpret_cast<uint2*>(y_bf16 + (int64_t)blockIdx.x * D)[i] =
          make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    if (y_f32) reinterpret_cast<float4*>(y_f32 + (int64_t)blockIdx.x * D)[i] = o;
  };
#pragma unroll
  for (int k = 0; k < kCache; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < D4) emit(i, v[k], __ldg(reinterpret_cast<const float4*>(w) + i));
  }
  for (int i = threadIdx.x + kCache * 256; i < D4; i += 256)
    emit(i, reinterpret_cast<const float4*>(xr)[i], __ldg(reinterpret_cast<const float4*>(w) + i));
}

// ------------------------------------------------------------------------------------------
// row softmax (reprogramming scores): one warp per row; bytes per row: 4n read + 2n written
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int64_t rows, int n,
                    float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * n;
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, sr[i]);
  mx = warp_max(mx) * scale;
  float sum = 0.0f;
  for (int i = lane; i < n; i += 32) sum += __expf(sr[i] * scale - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = p + row * n;
  for (int i = lane; i < n; i += 32) pr[i] = __float2bfloat16_rn(__expf(sr[i] * scale - mx) * inv);
}

// ------------------------------------------------------------------------------------------
// SwiGLU on a saved [rows, 2I] = [g | u] buffer (training path)
// ------------------------------------------------------------------------------------------
__global__ void swiglu_kernel(const __nv_bfloat16* __restrict__ gu, int64_t ldgu,
                              __nv_bfloat16* __restrict__ y, int64_t rows, int I) {
  const int I8 = I >> 3;
  const int64_t total = rows * I8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / I8;
    const int c = (int)(idx - r * I8) * 8;
    const uint4 g = *reinterpret_cast<const uint4*>(gu + r * ldgu + c);
    const uint4 u = *reinterpret_cast<const uint4*>(gu + r * ldgu + I + c);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g0 = bf16_lo(gw[q]), g1 = bf16_hi(gw[q]);
      const float u0 = bf16_lo(uw[q]), u1 = bf16_hi(uw[q]);
      o[q] = pack_bf16(g0 / (1.0f + __expf(-g0)) * u0, g1 / (1.0f + __expf(-g1)) * u1);
    }
    *reinterpret_cast<uint4*>(y + r * I + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------
// gate/up packing for the SWIGLU epilogue (one-off at weight load)
// ------------------------------------------------------------------------------------------
__global__ void pack_gate_up_kernel(const __nv_bfloat16* __restrict__ gate,
                                    const __nv_bfloat16* __restrict__ up,
                                    __nv_bfloat16* __restrict__ out, int I, int K) {
  // out row r: block = r/256, within = r%256; within<128 -> gate[block*128+within], else up[...]
  const int r = blockIdx.x;
  const int blk = r >> 8, within = r & 255;
  const int src = blk * 128 + (within & 127);
  const __nv_bfloat16* s = (within < 128 ? gate : up) + (int64_t)src * K;
  __nv_bfloat16* d = out + (int64_t)r * K;
  const bool valid = src < I;
  for (int i = threadIdx.x; i < K; i += blockDim.x) d[i] = valid ? s[i] : __float2bfloat16_rn(0.0f);
}

// ------------------------------------------------------------------------------------------
// prompt gather: one CTA per (b, l) row; bytes per row: 4*D read (+4*D wpe) + 4*D*rep written
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prompt_gather_kernel(const int32_t* __restrict__ ids, const float* __restrict__ emb,
                     const float* __restrict__ wpe, float* __restrict__ x, int rep, int Lp, int L,
                     int Lc, int D) {
  // Lc > 0 (shared-prefix layout): the first Lc prompt tokens are the same for every sample and are written
  // once, to rows [0, Lc) (taken from ids row 0); sample b then owns rows Lc + b*(L-Lc) + [0, L-Lc).
  const int Ls = L - Lc;
  int b, l;
  int64_t row0;          // output row of (b, replica 0, l)
  int64_t rep_stride;    // rows between replicas
  if ((int)blockIdx.x < Lc) {
    b = 0; l = blockIdx.x; row0 = l; rep_stride = 0; rep = 1;
  } else {
    const int j = blockIdx.x - Lc;
    b = j / Ls; l = Lc + (j - b * Ls);
    row0 = (int64_t)Lc + (int64_t)b * rep * Ls + (l - Lc); rep_stride = Ls;
  }
  const int D4 = D >> 2;
  const float4* src = nullptr;
  if (l < Lp) {
    // a negative id marks a prompt position that is not a token (a time-series example part, models/medtsllm.py:
    // 313-319): the row starts as zero (+ wpe) and the reprogramming out-projection accumulates onto it
    const int32_t id = ids[(int64_t)b * Lp + l];
    if (id >= 0) src = reinterpret_cast<const float4*>(emb + (int64_t)id * D);
  }
  const float4* pe = wpe ? reinterpret_cast<const float4*>(wpe + (int64_t)l * D) : nullptr;
  for (int i = threadIdx.x; i < D4; i += blockDim.x) {
    float4 v = src ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (pe) {
      const float4 q = pe[i];
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    for (int r = 0; r < rep; ++r)
      reinterpret_cast<float4*>(x + (row0 + r * rep_stride) * D)[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// eval-only output activations (models/medtsllm.py:251-259): tiny, in place
// ------------------------------------------------------------------------------------------
__global__ void sigmoid_kernel(float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = 1.0f / (1.0f + expf(-y[i]));
}
__global__ void softmax_lastdim_kernel(float* __restrict__ y, int64_t rows, int n) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    float* p = y + r * n;
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) mx = fmaxf(mx, p[i]);
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) sum += expf(p[i] - mx);
    const float inv = 1.0f / sum;
    for (int i = 0; i < n; ++i) p[i] = expf(p[i] - mx) * inv;
  }
}

static int grid_for(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace mts

using namespace mts;

extern "C" int mts_cast_f32_bf16(const float* in, uint16_t* out, int64_t n, mts_stream_t s) {
  if (!in || !out || n < 0) return set_error(MTS_ERR_INVALID_ARG, "mts_cast_f32_bf16: bad args");
  if (n == 0) return MTS_OK;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_cast_f32_bf16: pointers must be 16-byte aligned");
  cast_f32_bf16_kernel<<<grid_for(n, 256 * 8), 256, 0, (cudaStream_t)s>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out), n);
  count_launch();
  return check_launch("cast_f32_bf16_kernel");
}

extern "C" int mts_cast_bf16_f32(const uint16_t* in, float* out, int64_t n, mts_stream_t s) {
  if (!in || !out || n < 0) return set_error(MTS_ERR_INVALID_ARG, "mts_cast_bf16_f32: bad args");
  if (n == 0) return MTS_OK;
  cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)s>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), out, n);
  count_launch();
  return check_launch("cast_bf16_f32_kernel");
}

extern "C" int mts_transpose_f32_bf16(const float* in, uint16_t* out, int rows, int cols,
                                      mts_stream_t s) {
  if (!in || !out || rows <= 0 || cols <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_transpose_f32_bf16: bad args");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_to_bf16_kernel<float><<<grid, block, 0, (cudaStream_t)s>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out), rows, cols);
  count_launch();
  return check_launch("transpose_to_bf16_kernel<float>");
}

extern "C" int mts_transpose_bf16(const uint16_t* in, uint16_t* out, int rows, int cols,
                                  mts_stream_t s) {
  if (!in || !out || rows <= 0 || cols <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_transpose_bf16: bad args");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_to_bf16_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)s>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), rows, cols);
  count_launch();
  return check_launch("transpose_to_bf16_kernel<bf16>");
}

static int norm_common(bool ln, const float* x, int64_t ldx, const float* w, const float* b,
                       uint16_t* y_bf16, float* y_f32, int rows, int D, float eps, mts_stream_t s) {
  const char* name = ln ? "mts_layernorm" : "mts_rmsnorm";
  if (!x || !w || (ln && !b) || (!y_bf16 && !y_f32) || rows < 0 || D <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "%s: bad args", name);
  if ((D % 4) || (ldx % 4) || ldx < D || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (b && (reinterpret_cast<uintptr_t>(b) & 15)))
    return set_error(MTS_ERR_INVALID_ARG, "%s: D, ldx must be multiples of 4 and pointers 16-byte aligned", name);
  if (rows == 0) return MTS_OK;
  if (ln)
    LAUNCH_PDL(norm_rows_kernel<true>, rows, 256, 0, s,
        x, ldx, w, b, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, D, eps);
  else
    LAUNCH_PDL(norm_rows_kernel<false>, rows, 256, 0, s,
        x, ldx, w, nullptr, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, D, eps);
  count_launch();
  return check_launch(name);
}

extern "C" int mts_rmsnorm(const float* x, int64_t ldx, const float* w, uint16_t* y_bf16,
                           float* y_f32, int rows, int D, float eps, mts_stream_t s) {
  return norm_common(false, x, ldx, w, nullptr, y_bf16, y_f32, rows, D, eps, s);
}
extern "C" int mts_layernorm(const float* x, int64_t ldx, const float* w, const float* b,
                             uint16_t* y_bf16, float* y_f32, int rows, int D, float eps,
                             mts_stream_t s) {
  return norm_common(true, x, ldx, w, b, y_bf16, y_f32, rows, D, eps, s);
}

extern "C" int mts_softmax_rows(const float* sc, uint16_t* p, int64_t rows, int n, float scale,
                                mts_stream_t s) {
  if (!sc || !p || rows < 0 || n <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_softmax_rows: bad args");
  if (rows == 0) return MTS_OK;
  const int64_t blocks = (rows + 7) / 8;
  if (blocks > 0x7fffffffLL) return set_error(MTS_ERR_INVALID_ARG, "mts_softmax_rows: too many rows");
  softmax_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(
      sc, reinterpret_cast<__nv_bfloat16*>(p), rows, n, scale);
  count_launch();
  return check_launch("softmax_rows_kernel");
}

extern "C" int mts_swiglu(const uint16_t* gu, int64_t ldgu, uint16_t* y, int64_t rows, int I,
                          mts_stream_t s) {
  if (!gu || !y || rows < 0 || I <= 0 || (I % 8) || (ldgu % 8) || ldgu < 2 * (int64_t)I)
    return set_error(MTS_ERR_INVALID_ARG, "mts_swiglu: bad args (I, ldgu multiples of 8)");
  if (rows == 0) return MTS_OK;
  swiglu_kernel<<<grid_for(rows * (I / 8), 256), 256, 0, (cudaStream_t)s>>>(
      reinterpret_cast<const __nv_bfloat16*>(gu), ldgu, reinterpret_cast<__nv_bfloat16*>(y), rows, I);
  count_launch();
  return check_launch("swiglu_kernel");
}

extern "C" int mts_pack_gate_up(const uint16_t* gate, const uint16_t* up, uint16_t* out, int I,
                                int K, mts_stream_t s) {
  if (!gate || !up || !out || I <= 0 || K <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_pack_gate_up: bad args");
  const int rows = ((I + 127) / 128) * 256;
  pack_gate_up_kernel<<<rows, 256, 0, (cudaStream_t)s>>>(
      reinterpret_cast<const __nv_bfloat16*>(gate), reinterpret_cast<const __nv_bfloat16*>(up),
      reinterpret_cast<__nv_bfloat16*>(out), I, K);
  count_launch();
  return check_launch("pack_gate_up_kernel");
}

extern "C" int mts_prompt_gather(const int32_t* ids, const float* emb, const float* wpe, float* x,
                                 int B, int rep, int Lp, int L, int D, mts_stream_t s) {
  if (!emb || !x || B <= 0 || rep <= 0 || Lp < 0 || L < Lp || L <= 0 || D <= 0 || (Lp > 0 && !ids))
    return set_error(MTS_ERR_INVALID_ARG, "mts_prompt_gather: bad args");
  if ((D % 4) || (reinterpret_cast<uintptr_t>(emb) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (wpe && (reinterpret_cast<uintptr_t>(wpe) & 15)))
    return set_error(MTS_ERR_INVALID_ARG, "mts_prompt_gather: D % 4 and 16-byte alignment required");
  prompt_gather_kernel<<<B * L, 256, 0, (cudaStream_t)s>>>(ids, emb, wpe, x, rep, Lp, L, 0, D);
  count_launch();
  return check_launch("prompt_gather_kernel");
}

extern "C" int mts_prompt_gather_shared(const int32_t* ids, const float* emb, const float* wpe, float* x,
                                        int B, int rep, int Lp, int L, int Lc, int D, mts_stream_t s) {
  if (!emb || !x || !ids || B <= 0 || rep <= 0 || Lp <= 0 || L < Lp || Lc <= 0 || Lc > Lp || Lc >= L || D <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_prompt_gather_shared: bad args (0 < Lc <= Lp <= L, Lc < L)");
  if ((D % 4) || (reinterpret_cast<uintptr_t>(emb) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (wpe && (reinterpret_cast<uintptr_t>(wpe) & 15)))
    return set_error(MTS_ERR_INVALID_ARG, "mts_prompt_gather_shared: D % 4 and 16-byte alignment required");
  prompt_gather_kernel<<<Lc + B * (L - Lc), 256, 0, (cudaStream_t)s>>>(ids, emb, wpe, x, rep, Lp, L, Lc, D);
  count_launch();
  return check_launch("prompt_gather_kernel");
}

extern "C" int mts_sigmoid(float* y, int64_t n, mts_stream_t s) {
  if (!y || n < 0) return set_error(MTS_ERR_INVALID_ARG, "mts_sigmoid: bad args");
  if (n == 0) return MTS_OK;
  sigmoid_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)s>>>(y, n);
  count_launch();
  return check_launch("sigmoid_kernel");
}

extern "C" int mts_softmax_lastdim(float* y, int64_t rows, int n, mts_stream_t s) {
  if (!y || rows < 0 || n <= 0) return set_error(MTS_ERR_INVALID_ARG, "mts_softmax_lastdim: bad args");
  if (rows == 0) return MTS_OK;
  softmax_lastdim_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)s>>>(y, rows, n);
  count_launch();
  return check_launch("softmax_lastdim_kernel");
}

// ------------------------------------------------------------------------------------------
// Dropout (training only; models/medtsllm.py:93-94 -> PatchEmbedding.dropout, embed.py:183,197, and the
// reprogramming attention dropout, :587).  Counter-based: the keep/drop decision of element i is a pure
// function of (seed, i), so the backward re-creates the forward's mask by calling the same kernel on the
// gradient with the same seed.  y = keep ? x / (1 - p) : 0.
// ------------------------------------------------------------------------------------------
namespace mts {
// (mix32 / dropout_keep: ptx.cuh — shared with the attention and GEMM-epilogue dropouts)
__global__ void dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n,
                                    uint32_t thresh, float scale, uint64_t seed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool keep = mix32(seed * 0xD1342543DE82EF95ull + (uint64_t)i) >= thresh;
    y[i] = keep ? __float2bfloat16_rn(__bfloat162float(x[i]) * scale) : __float2bfloat16_rn(0.f);
  }
}
__global__ void dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, uint32_t thresh,
                                   float scale, uint64_t seed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool keep = mix32(seed * 0xD1342543DE82EF95ull + (uint64_t)i) >= thresh;
    y[i] = keep ? x[i] * scale : 0.f;
  }
}
}  // namespace mts

extern "C" int mts_dropout(const void* x, void* y, int dtype, int64_t n, float p, uint64_t seed, mts_stream_t s) {
  if (!x || !y || n < 0 || !(p >= 0.f && p < 1.f)) return set_error(MTS_ERR_INVALID_ARG, "mts_dropout: bad args");
  if (n == 0) return MTS_OK;
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  const float scale = 1.0f / (1.0f - p);
  if (dtype == MTS_BF16)
    dropout_bf16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)s>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, thresh, scale, seed);
  else if (dtype == MTS_F32)
    dropout_f32_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)s>>>(static_cast<const float*>(x),
                                                                     static_cast<float*>(y), n, thresh, scale, seed);
  else
    return set_error(MTS_ERR_INVALID_ARG, "mts_dropout: bad dtype");
  count_launch();
  return check_launch("dropout_kernel");
}
