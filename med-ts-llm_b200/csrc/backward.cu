// backward.cu — row / element-wise kernels of the training path (dgrad through the frozen backbone
// and the adapter gradients).  The GEMM-shaped parts of the backward are the same tcgen05 NT kernel
// (gemm_tcgen05.cu) fed with transposed operands; attention backward is attention_bwd.cu.
//
// All HBM-bound: one pass over the operands (second reads of a row hit L1/L2), 16-byte vectors.
#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

__device__ __forceinline__ float bw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bw_block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = bw_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.0f;
  return bw_warp_sum(t);
}

// ------------------------------------------------------------------------------------------
// RMSNorm / LayerNorm backward w.r.t. the input (weights are frozen):  dx (+)= J^T (w * dy)
//   RMSNorm:   y = w * x * r,  r = rsqrt(mean(x^2)+eps):  dx = r*g - x * r^3 * mean(g*x)
//   LayerNorm: xh = (x-mu)*r:                            dx = r*(g - mean(g) - xh*mean(g*xh))
// ------------------------------------------------------------------------------------------
template <bool kLayerNorm>
__global__ void __launch_bounds__(256)
norm_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                const __nv_bfloat16* __restrict__ dy, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16,
                int D, float eps, int accumulate) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* xr = x + row * ldx;
  const __nv_bfloat16* dyr = dy + row * D;
  float* dxr = dx + row * D;
  const int D4 = D >> 2;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < D4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(xr)[i];
    s += (v.x + v.y) + (v.z + v.w);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  float mean = 0.f, rstd;
  if (kLayerNorm) {
    mean = bw_block_sum(s, red) / D;
    float vs = 0.f;
    for (int i = threadIdx.x; i < D4; i += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(xr)[i];
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      vs += a * a + b * b + c * c + d * d;
    }
    rstd = rsqrtf(bw_block_sum(vs, red) / D + eps);
  } else {
    rstd = rsqrtf(bw_block_sum(ss, red) / D + eps);
  }
  // reductions over g = w*dy
  float sg = 0.f, sgx = 0.f;
  for (int i = threadIdx.x; i < D4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(xr)[i];
    const float4 g4 = reinterpret_cast<const float4*>(w)[i];
    const uint2 d2 = reinterpret_cast<const uint2*>(dyr)[i];
    const float g0 = g4.x * bf16_lo(d2.x), g1 = g4.y * bf16_hi(d2.x);
    const float g2 = g4.z * bf16_lo(d2.y), g3 = g4.w * bf16_hi(d2.y);
    sg += (g0 + g1) + (g2 + g3);
    sgx += g0 * (v.x - mean) + g1 * (v.y - mean) + g2 * (v.z - mean) + g3 * (v.w - mean);
  }
  const float mg = kLayerNorm ? bw_block_sum(sg, red) / D : 0.f;
  const float mgx = bw_block_sum(sgx, red) / D * rstd;  // mean(g * xh) with xh = (x-mean)*rstd
  for (int i = threadIdx.x; i < D4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(xr)[i];
    const float4 g4 = reinterpret_cast<const float4*>(w)[i];
    const uint2 d2 = reinterpret_cast<const uint2*>(dyr)[i];
    const float g[4] = {g4.x * bf16_lo(d2.x), g4.y * bf16_hi(d2.x), g4.z * bf16_lo(d2.y), g4.w * bf16_hi(d2.y)};
    const float xv[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float xh = (xv[q] - mean) * rstd;
      o[q] = rstd * (g[q] - mg - xh * mgx);
    }
    float4 out = make_float4(o[0], o[1], o[2], o[3]);
    if (accumulate) {
      const float4 prev = reinterpret_cast<const float4*>(dxr)[i];
      out.x += prev.x; out.y += prev.y; out.z += prev.z; out.w += prev.w;
    }
    reinterpret_cast<float4*>(dxr)[i] = out;
    if (dx_bf16)   // bf16 copy of the updated gradient: the A operand of the next dgrad GEMM
      reinterpret_cast<uint2*>(dx_bf16 + row * D)[i] = make_uint2(pack_bf16(out.x, out.y), pack_bf16(out.z, out.w));
  }
}

// ------------------------------------------------------------------------------------------
// SwiGLU forward / backward on a [rows, ld] buffer of pre-activations.  Column j of the
// activation reads gate column gc(j) = (j/blk)*2*blk + j%blk and up column gc(j)+blk:
// blk = 128 for the packed layout of the fused-epilogue weights, blk = I for plain [g | u].
// ------------------------------------------------------------------------------------------
__global__ void swiglu_blk_kernel(const __nv_bfloat16* __restrict__ gu, int64_t ld,
                                  __nv_bfloat16* __restrict__ y, int64_t rows, int I, int blk) {
  const int I8 = I >> 3;
  const int64_t total = rows * I8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / I8;
    const int j = (int)(idx - r * I8) * 8;
    const int gc = (j / blk) * 2 * blk + (j % blk);
    const uint4 g = *reinterpret_cast<const uint4*>(gu + r * ld + gc);
    const uint4 u = *reinterpret_cast<const uint4*>(gu + r * ld + gc + blk);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g0 = bf16_lo(gw[q]), g1 = bf16_hi(gw[q]);
      o[q] = pack_bf16(g0 / (1.f + __expf(-g0)) * bf16_lo(uw[q]), g1 / (1.f + __expf(-g1)) * bf16_hi(uw[q]));
    }
    *reinterpret_cast<uint4*>(y + r * I + j) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, int64_t ld,
                                  const __nv_bfloat16* __restrict__ dact, __nv_bfloat16* __restrict__ dgu,
                                  int64_t rows, int I, int blk) {
  pdl_wait();
  pdl_trigger();
  const int I8 = I >> 3;
  const int64_t total = rows * I8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / I8;
    const int j = (int)(idx - r * I8) * 8;
    const int gc = (j / blk) * 2 * blk + (j % blk);
    const uint4 g = *reinterpret_cast<const uint4*>(gu + r * ld + gc);
    const uint4 u = *reinterpret_cast<const uint4*>(gu + r * ld + gc + blk);
    const uint4 d = *reinterpret_cast<const uint4*>(dact + r * I + j);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w}, dw[4] = {d.x, d.y, d.z, d.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float dg[2], du[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float gv = h ? bf16_hi(gw[q]) : bf16_lo(gw[q]);
        const float uv = h ? bf16_hi(uw[q]) : bf16_lo(uw[q]);
        const float dv = h ? bf16_hi(dw[q]) : bf16_lo(dw[q]);
        const float sg = 1.f / (1.f + __expf(-gv));
        du[h] = dv * gv * sg;                                  // d/du = silu(g)
        dg[h] = dv * uv * sg * (1.f + gv * (1.f - sg));        // d/dg = u * silu'(g)
      }
      og[q] = pack_bf16(dg[0], dg[1]);
      ou[q] = pack_bf16(du[0], du[1]);
    }
    *reinterpret_cast<uint4*>(dgu + r * ld + gc) = make_uint4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<uint4*>(dgu + r * ld + gc + blk) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

// ------------------------------------------------------------------------------------------
// gelu_new forward / backward on saved pre-activations (GPT-2 training path)
// ------------------------------------------------------------------------------------------
__global__ void gelu_new_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dact,
                                __nv_bfloat16* __restrict__ out, int64_t n) {
  pdl_wait();
  pdl_trigger();
  // dact == nullptr: out = gelu_new(pre); else out = dact * gelu_new'(pre)
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n;
       i += (int64_t)gridDim.x * blockDim.x * 2) {
    const uint32_t pw = *reinterpret_cast<const uint32_t*>(pre + i);
    const uint32_t dw = dact ? *reinterpret_cast<const uint32_t*>(dact + i) : 0u;
    float o[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float x = h ? bf16_hi(pw) : bf16_lo(pw);
      const float c = 0.7978845608028654f;
      const float t = tanhf(c * (x + 0.044715f * x * x * x));
      if (dact) {
        const float d = h ? bf16_hi(dw) : bf16_lo(dw);
        o[h] = d * (0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * c * (1.f + 3.f * 0.044715f * x * x));
      } else {
        o[h] = 0.5f * x * (1.f + t);
      }
    }
    *reinterpret_cast<uint32_t*>(out + i) = pack_bf16(o[0], o[1]);
  }
}

// ------------------------------------------------------------------------------------------
// softmax backward (reprogramming scores): ds = scale * p * (dp - sum(dp*p)); one warp per row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                        __nv_bfloat16* __restrict__ ds, int64_t rows, int n, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* pr = p + row * n;
  const float* dr = dp + row * n;
  float dot = 0.f;
  for (int i = lane; i < n; i += 32) dot += __bfloat162float(pr[i]) * dr[i];
  dot = bw_warp_sum(dot);
  __nv_bfloat16* o = ds + row * n;
  for (int i = lane; i < n; i += 32)
    o[i] = __float2bfloat16_rn(scale * __bfloat162float(pr[i]) * (dr[i] - dot));
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients): out[c] = sum_r x[r, c]
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int64_t ld, float* __restrict__ out, int rows, int cols) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < cols)
    for (int r = w; r < rows; r += 8) acc += static_cast<float>(x[(int64_t)r * ld + c]);
  part[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
    out[c] = s;
  }
}

// row sums in fp32 (mapping-layer bias gradient = row sums of dSource): one warp per row
__global__ void __launch_bounds__(256)
rowsum_kernel(const float* __restrict__ x, int64_t ld, float* __restrict__ out, int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * ld;
  float acc = 0.f;
  for (int i = lane; i < cols; i += 32) acc += xr[i];
  acc = bw_warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

// ------------------------------------------------------------------------------------------
// general strided transpose with cast to bf16:
//   out[c * ld_out + (b*rows + r)] = in[b*in_bs + r*ld_in + c]      (b < batch, r < rows, c < cols)
// ------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void transpose_strided_kernel(const TIn* __restrict__ in, int64_t ld_in, int64_t in_bs,
                                         __nv_bfloat16* __restrict__ out, int64_t ld_out, int rows,
                                         int cols) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const TIn* src = in + (int64_t)b * in_bs;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? static_cast<float>(src[(int64_t)r * ld_in + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols)
      out[(int64_t)c * ld_out + (int64_t)b * rows + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

// strided 3-D gather + cast: out[b*out_bs + r*ld_out + c] = bf16(in[b*in_bs + r*ld_in + c])
__global__ void cast_rows_kernel(const float* __restrict__ in, int64_t ld_in, int64_t in_bs,
                                 __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t out_bs, int batch,
                                 int rows, int cols, int vec_ok) {
  if (vec_ok) {
    const int c4n = cols >> 2;
    const int64_t total = (int64_t)batch * rows * c4n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
      const int c4 = (int)(idx % c4n);
      const int64_t br = idx / c4n;
      const int r = (int)(br % rows);
      const int b = (int)(br / rows);
      const float4 v = *reinterpret_cast<const float4*>(in + (int64_t)b * in_bs + (int64_t)r * ld_in + c4 * 4);
      *reinterpret_cast<uint2*>(out + (int64_t)b * out_bs + (int64_t)r * ld_out + c4 * 4) =
          make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
  } else {
    const int64_t total = (int64_t)batch * rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
      const int c = (int)(idx % cols);
      const int64_t br = idx / cols;
      const int r = (int)(br % rows);
      const int b = (int)(br / rows);
      out[(int64_t)b * out_bs + (int64_t)r * ld_out + c] =
          __float2bfloat16_rn(in[(int64_t)b * in_bs + (int64_t)r * ld_in + c]);
    }
  }
}

// d(denorm): out[b,t,c] = dy[b,t,c] * stdev[b,c]
__global__ void scale_by_std_kernel(const float* __restrict__ dy, const float* __restrict__ stdev,
                                    float* __restrict__ out, int64_t total, int T, int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t b = i / ((int64_t)T * C);
    out[i] = dy[i] * stdev[b * C + c];
  }
}

static int bw_grid(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace mts

using namespace mts;

static int norm_bwd_common(bool ln, const float* x, int64_t ldx, const float* w, const uint16_t* dy,
                           float* dx, uint16_t* dx_bf16, int rows, int D, float eps, int accumulate, mts_stream_t s) {
  const char* name = ln ? "mts_layernorm_bwd" : "mts_rmsnorm_bwd";
  if (!x || !w || !dy || !dx || rows < 0 || D <= 0 || (D % 4) || (ldx % 4) || ldx < D)
    return set_error(MTS_ERR_INVALID_ARG, "%s: bad args", name);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
      (reinterpret_cast<uintptr_t>(dy) & 7) || (reinterpret_cast<uintptr_t>(dx) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "%s: misaligned pointer", name);
  if (rows == 0) return MTS_OK;
  if (ln)
    LAUNCH_PDL(norm_bwd_kernel<true>, rows, 256, 0, (cudaStream_t)s,
      
        x, ldx, w, reinterpret_cast<const __nv_bfloat16*>(dy), dx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), D, eps,
        accumulate);
  else
    LAUNCH_PDL(norm_bwd_kernel<false>, rows, 256, 0, (cudaStream_t)s,
      
        x, ldx, w, reinterpret_cast<const __nv_bfloat16*>(dy), dx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), D, eps,
        accumulate);
  count_launch();
  return check_launch(name);
}

extern "C" int mts_rmsnorm_bwd(const float* x, int64_t ldx, const float* w, const uint16_t* dy,
                               float* dx, uint16_t* dx_bf16, int rows, int D, float eps, int accumulate,
                               mts_stream_t s) {
  return norm_bwd_common(false, x, ldx, w, dy, dx, dx_bf16, rows, D, eps, accumulate, s);
}
extern "C" int mts_layernorm_bwd(const float* x, int64_t ldx, const float* w, const uint16_t* dy,
                                 float* dx, uint16_t* dx_bf16, int rows, int D, float eps, int accumulate,
                                 mts_stream_t s) {
  return norm_bwd_common(true, x, ldx, w, dy, dx, dx_bf16, rows, D, eps, accumulate, s);
}

extern "C" int mts_swiglu_blk(const uint16_t* gu, int64_t ld, uint16_t* y, int64_t rows, int I, int blk,
                              mts_stream_t s) {
  if (!gu || !y || rows < 0 || I <= 0 || (I % 8) || (ld % 8) || blk <= 0 || (blk % 8) || (I % blk) ||
      ld < 2 * (int64_t)I)
    return set_error(MTS_ERR_INVALID_ARG, "mts_swiglu_blk: bad args");
  if (rows == 0) return MTS_OK;
  swiglu_blk_kernel<<<bw_grid(rows * (I / 8), 256), 256, 0, (cudaStream_t)s>>>(
      reinterpret_cast<const __nv_bfloat16*>(gu), ld, reinterpret_cast<__nv_bfloat16*>(y), rows, I, blk);
  count_launch();
  return check_launch("swiglu_blk_kernel");
}

extern "C" int mts_swiglu_bwd(const uint16_t* gu, int64_t ld, const uint16_t* dact, uint16_t* dgu,
                              int64_t rows, int I, int blk, mts_stream_t s) {
  if (!gu || !dact || !dgu || rows < 0 || I <= 0 || (I % 8) || (ld % 8) || blk <= 0 || (blk % 8) ||
      (I % blk) || ld < 2 * (int64_t)I)
    return set_error(MTS_ERR_INVALID_ARG, "mts_swiglu_bwd: bad args");
  if (rows == 0) return MTS_OK;
  LAUNCH_PDL(swiglu_bwd_kernel, bw_grid(rows * (I / 8), 256), 256, 0, (cudaStream_t)s,
      
      reinterpret_cast<const __nv_bfloat16*>(gu), ld, reinterpret_cast<const __nv_bfloat16*>(dact),
      reinterpret_cast<__nv_bfloat16*>(dgu), rows, I, blk);
  count_launch();
  return check_launch("swiglu_bwd_kernel");
}

extern "C" int mts_gelu_new(const uint16_t* pre, const uint16_t* dact, uint16_t* out, int64_t n,
                            mts_stream_t s) {
  if (!pre || !out || n < 0 || (n % 2))
    return set_error(MTS_ERR_INVALID_ARG, "mts_gelu_new: bad args (n must be even)");
  if (n == 0) return MTS_OK;
  LAUNCH_PDL(gelu_new_kernel, bw_grid(n / 2, 256), 256, 0, (cudaStream_t)s,
      
      reinterpret_cast<const __nv_bfloat16*>(pre), reinterpret_cast<const __nv_bfloat16*>(dact),
      reinterpret_cast<__nv_bfloat16*>(out), n);
  count_launch();
  return check_launch("gelu_new_kernel");
}

extern "C" int mts_softmax_bwd_rows(const uint16_t* p, const float* dp, uint16_t* ds, int64_t rows,
                                    int n, float scale, mts_stream_t s) {
  if (!p || !dp || !ds || rows < 0 || n <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_softmax_bwd_rows: bad args");
  if (rows == 0) return MTS_OK;
  const int64_t blocks = (rows + 7) / 8;
  if (blocks > 0x7fffffffLL) return set_error(MTS_ERR_INVALID_ARG, "mts_softmax_bwd_rows: too many rows");
  softmax_bwd_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(
      reinterpret_cast<const __nv_bfloat16*>(p), dp, reinterpret_cast<__nv_bfloat16*>(ds), rows, n, scale);
  count_launch();
  return check_launch("softmax_bwd_rows_kernel");
}

// ------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm parameter gradients (GPT4TS trains the GPT-2 LayerNorms, models/gpt4ts.py:47-53):
//   dgamma[d] = sum_r dy[r, d] * xhat[r, d],  dbeta[d] = sum_r dy[r, d],  xhat recomputed from x.
// Stage 1 (this kernel): one CTA per 32 rows writes its partial sums [2*D] (row statistics by one warp per row,
// then one thread per column); stage 2 = mts_colsum over the partials.  Deterministic.
// ------------------------------------------------------------------------------------------
namespace mts {
template <bool kLayerNorm>
__global__ void __launch_bounds__(256)
norm_wgrad_partial_kernel(const float* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ dy,
                          float* __restrict__ partial, int rows, int D, float eps) {
  __shared__ float s_mean[32], s_rstd[32];
  const int r0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < 32; j += 8) {
    const int r = r0 + j;
    float mean = 0.f, rstd = 0.f;
    if (r < rows) {
      const float* xr = x + (int64_t)r * ldx;
      float sum = 0.f;
      if (kLayerNorm) {
        for (int i = lane; i < D; i += 32) sum += xr[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mean = sum / D;
      }
      float ss = 0.f;
      for (int i = lane; i < D; i += 32) { const float v = xr[i] - mean; ss += v * v; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      rstd = rsqrtf(ss / D + eps);
    }
    if (lane == 0) { s_mean[j] = mean; s_rstd[j] = rstd; }
  }
  __syncthreads();
  const int nr = min(32, rows - r0);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float g = 0.f, b = 0.f;
    for (int j = 0; j < nr; ++j) {
      const float dyv = __bfloat162float(dy[(int64_t)(r0 + j) * D + d]);
      g += dyv * (x[(int64_t)(r0 + j) * ldx + d] - s_mean[j]) * s_rstd[j];
      b += dyv;
    }
    partial[(int64_t)blockIdx.x * 2 * D + d] = g;
    partial[(int64_t)blockIdx.x * 2 * D + D + d] = b;
  }
}
}  // namespace mts

extern "C" int mts_norm_wgrad_partial(const float* x, int64_t ldx, const uint16_t* dy, float* partial, int rows, int D,
                                      float eps, int layernorm, mts_stream_t s) {
  if (!x || !dy || !partial || rows <= 0 || D <= 0 || ldx < D)
    return set_error(MTS_ERR_INVALID_ARG, "mts_norm_wgrad_partial: bad args");
  const int grid = (rows + 31) / 32;
  if (layernorm)
    norm_wgrad_partial_kernel<true><<<grid, 256, 0, (cudaStream_t)s>>>(x, ldx, reinterpret_cast<const __nv_bfloat16*>(dy),
                                                                      partial, rows, D, eps);
  else
    norm_wgrad_partial_kernel<false><<<grid, 256, 0, (cudaStream_t)s>>>(x, ldx, reinterpret_cast<const __nv_bfloat16*>(dy),
                                                                       partial, rows, D, eps);
  count_launch();
  return check_launch("norm_wgrad_partial_kernel");
}

extern "C" int mts_colsum(const void* x, int dtype, int64_t ld, float* out, int rows, int cols,
                          mts_stream_t s) {
  if (!x || !out || rows <= 0 || cols <= 0 || ld < cols)
    return set_error(MTS_ERR_INVALID_ARG, "mts_colsum: bad args");
  const int grid = (cols + 31) / 32;
  if (dtype == MTS_F32)
    colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(static_cast<const float*>(x), ld, out, rows, cols);
  else if (dtype == MTS_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)s>>>(
        static_cast<const __nv_bfloat16*>(x), ld, out, rows, cols);
  else
    return set_error(MTS_ERR_INVALID_ARG, "mts_colsum: bad dtype");
  count_launch();
  return check_launch("colsum_kernel");
}

extern "C" int mts_transpose_strided(const void* in, int dtype, int64_t ld_in, int64_t in_batch_stride,
                                     uint16_t* out, int64_t ld_out, int batch, int rows, int cols,
                                     mts_stream_t s) {
  if (!in || !out || batch <= 0 || rows <= 0 || cols <= 0 || ld_in < cols ||
      ld_out < (int64_t)batch * rows || batch > 65535)
    return set_error(MTS_ERR_INVALID_ARG, "mts_transpose_strided: bad args");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
  if (grid.y > 65535) return set_error(MTS_ERR_INVALID_ARG, "mts_transpose_strided: too many rows");
  if (dtype == MTS_F32)
    transpose_strided_kernel<float><<<grid, block, 0, (cudaStream_t)s>>>(
        static_cast<const float*>(in), ld_in, in_batch_stride, reinterpret_cast<__nv_bfloat16*>(out),
        ld_out, rows, cols);
  else if (dtype == MTS_BF16)
    transpose_strided_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)s>>>(
        static_cast<const __nv_bfloat16*>(in), ld_in, in_batch_stride,
        reinterpret_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  else
    return set_error(MTS_ERR_INVALID_ARG, "mts_transpose_strided: bad dtype");
  count_launch();
  return check_launch("transpose_strided_kernel");
}

extern "C" int mts_cast_rows_f32_bf16(const float* in, int64_t ld_in, int64_t in_batch_stride,
                                      uint16_t* out, int64_t ld_out, int64_t out_batch_stride, int batch,
                                      int rows, int cols, mts_stream_t s) {
  if (!in || !out || batch <= 0 || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < cols)
    return set_error(MTS_ERR_INVALID_ARG, "mts_cast_rows_f32_bf16: bad args");
  if (out_batch_stride == 0) out_batch_stride = (int64_t)rows * ld_out;
  const int vec_ok = !(cols % 4) && !(ld_in % 4) && !(in_batch_stride % 4) && !(ld_out % 4) && !(out_batch_stride % 4) &&
                     !(reinterpret_cast<uintptr_t>(in) & 15) && !(reinterpret_cast<uintptr_t>(out) & 7);
  const int64_t work = (int64_t)batch * rows * (vec_ok ? cols / 4 : cols);
  cast_rows_kernel<<<bw_grid(work, 256), 256, 0, (cudaStream_t)s>>>(
      in, ld_in, in_batch_stride, reinterpret_cast<__nv_bfloat16*>(out), ld_out, out_batch_stride, batch, rows, cols,
      vec_ok);
  count_launch();
  return check_launch("cast_rows_kernel");
}

extern "C" int mts_revin_denorm_bwd(const float* dy, const float* stdev, float* out, int B, int T, int C,
                                    mts_stream_t s) {
  if (!dy || !stdev || !out || B <= 0 || T <= 0 || C <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_revin_denorm_bwd: bad args");
  const int64_t total = (int64_t)B * T * C;
  scale_by_std_kernel<<<bw_grid(total, 256), 256, 0, (cudaStream_t)s>>>(dy, stdev, out, total, T, C);
  count_launch();
  return check_launch("scale_by_std_kernel");
}

extern "C" int mts_rowsum_f32(const float* x, int64_t ld, float* out, int rows, int cols, mts_stream_t s) {
  if (!x || !out || rows <= 0 || cols <= 0 || ld < cols)
    return set_error(MTS_ERR_INVALID_ARG, "mts_rowsum_f32: bad args");
  rowsum_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(x, ld, out, rows, cols);
  count_launch();
  return check_launch("rowsum_kernel");
}

// ------------------------------------------------------------------------------------------
// Covariate merges over the feature axis C (models/medtsllm.py:284-295, 369-377).  All tiny,
// HBM-bound; one element of the merged output per thread.
// ------------------------------------------------------------------------------------------
namespace mts {

// out[b*out_bs + r] = sum_c w[c] * in[(b*C + c)*R + r] + bias   (w == NULL: 1/C; bias may be NULL)
__global__ void group_reduce_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ out, int64_t out_bs,
                                    int B, int C, int64_t R, int accumulate) {
  const int64_t total = (int64_t)B * R;
  const float b0 = bias ? bias[0] : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / R, r = i - b * R;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += (w ? w[c] : 1.0f / C) * in[(b * C + c) * R + r];
    out[b * out_bs + r] = acc + b0 + (accumulate ? out[b * out_bs + r] : 0.f);
  }
}

// din[(b*C + c)*R + r] = w[c] * dout[b*dout_bs + r]     (w == NULL: 1/C)
__global__ void group_broadcast_kernel(const float* __restrict__ dout, int64_t dout_bs, const float* __restrict__ w,
                                       float* __restrict__ din, int B, int C, int64_t R) {
  const int64_t total = (int64_t)B * C * R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % R;
    const int64_t bc = i / R;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    din[i] = (w ? w[c] : 1.0f / C) * dout[b * dout_bs + r];
  }
}

// dw[c] = sum_{b,r} dout[b*dout_bs + r] * in[(b*C + c)*R + r];  dbias = sum dout   (one CTA per c, +1 for the bias)
__global__ void __launch_bounds__(256)
group_weight_grad_kernel(const float* __restrict__ dout, int64_t dout_bs, const float* __restrict__ in,
                         float* __restrict__ dw, float* __restrict__ dbias, int B, int C, int64_t R) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float acc = 0.f;
  const int64_t total = (int64_t)B * R;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int64_t b = i / R, r = i - b * R;
    const float d = dout[b * dout_bs + r];
    acc += (c < C) ? d * in[(b * C + c) * R + r] : d;
  }
  acc = bw_block_sum(acc, red);
  if (threadIdx.x == 0) { if (c < C) dw[c] = acc; else dbias[0] = acc; }
}

// merge-end: y[b,p,o] = sum_{o2,c} W[o, o2*C + c] * h[((b*C + c)*P + p)*O + o2] + bias[o]   (Linear(C*O -> O))
__global__ void merge_end_kernel(const float* __restrict__ h, const float* __restrict__ W, const float* __restrict__ bias,
                                 float* __restrict__ y, int B, int C, int P, int O) {
  const int64_t total = (int64_t)B * P * O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const int64_t bp = i / O;
    const int p = (int)(bp % P);
    const int64_t b = bp / P;
    float acc = bias[o];
    for (int o2 = 0; o2 < O; ++o2)
      for (int c = 0; c < C; ++c) acc += W[(int64_t)o * O * C + o2 * C + c] * h[((b * C + c) * P + p) * O + o2];
    y[i] = acc;
  }
}
// dh[((b*C + c)*P + p)*O + o2] = sum_o dy[b,p,o] * W[o, o2*C + c]
__global__ void merge_end_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                           float* __restrict__ dh, int B, int C, int P, int O) {
  const int64_t total = (int64_t)B * C * P * O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o2 = (int)(i % O);
    int64_t t = i / O;
    const int p = (int)(t % P); t /= P;
    const int c = (int)(t % C);
    const int64_t b = t / C;
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc += dy[(b * P + p) * O + o] * W[(int64_t)o * O * C + o2 * C + c];
    dh[i] = acc;
  }
}
// dW[o, o2*C + c] = sum_{b,p} dy[b,p,o] * h[b,c,p,o2];  dbias[o] = sum_{b,p} dy[b,p,o]   (one CTA per W entry / bias)
__global__ void __launch_bounds__(128)
merge_end_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ h, float* __restrict__ dW,
                            float* __restrict__ dbias, int B, int C, int P, int O) {
  __shared__ float red[32];
  const int e = blockIdx.x;               // < O*O*C: weight entry; >= : bias entry
  const int nW = O * O * C;
  float acc = 0.f;
  if (e < nW) {
    const int o = e / (O * C), o2 = (e / C) % O, c = e % C;
    for (int64_t i = threadIdx.x; i < (int64_t)B * P; i += blockDim.x) {
      const int64_t b = i / P; const int p = (int)(i - b * P);
      acc += dy[(b * P + p) * O + o] * h[((b * C + c) * P + p) * O + o2];
    }
  } else {
    const int o = e - nW;
    for (int64_t i = threadIdx.x; i < (int64_t)B * P; i += blockDim.x) acc += dy[i * O + o];
  }
  acc = bw_block_sum(acc, red);
  if (threadIdx.x == 0) { if (e < nW) dW[e] = acc; else dbias[e - nW] = acc; }
}

}  // namespace mts

extern "C" int mts_group_reduce(const float* in, const float* w, const float* bias, float* out, int64_t out_bs,
                                int B, int C, int64_t R, int accumulate, mts_stream_t s) {
  if (!in || !out || B <= 0 || C <= 0 || R <= 0 || out_bs < R)
    return set_error(MTS_ERR_INVALID_ARG, "mts_group_reduce: bad args");
  group_reduce_kernel<<<bw_grid((int64_t)B * R, 256), 256, 0, (cudaStream_t)s>>>(in, w, bias, out, out_bs, B, C, R, accumulate);
  count_launch();
  return check_launch("group_reduce_kernel");
}

extern "C" int mts_group_reduce_bwd(const float* dout, int64_t dout_bs, const float* w, const float* in, float* din,
                                    float* dw, float* dbias, int B, int C, int64_t R, mts_stream_t s) {
  if (!dout || !din || B <= 0 || C <= 0 || R <= 0 || dout_bs < R || ((dw != nullptr) != (in != nullptr)) ||
      (dw && !dbias))
    return set_error(MTS_ERR_INVALID_ARG, "mts_group_reduce_bwd: bad args");
  group_broadcast_kernel<<<bw_grid((int64_t)B * C * R, 256), 256, 0, (cudaStream_t)s>>>(dout, dout_bs, w, din, B, C, R);
  count_launch();
  int rc = check_launch("group_broadcast_kernel");
  if (rc || !dw) return rc;
  group_weight_grad_kernel<<<C + 1, 256, 0, (cudaStream_t)s>>>(dout, dout_bs, in, dw, dbias, B, C, R);
  count_launch();
  return check_launch("group_weight_grad_kernel");
}

extern "C" int mts_merge_end(const float* h, const float* W, const float* bias, float* y, int B, int C, int P, int O,
                             mts_stream_t s) {
  if (!h || !W || !bias || !y || B <= 0 || C <= 0 || P <= 0 || O <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_merge_end: bad args");
  merge_end_kernel<<<bw_grid((int64_t)B * P * O, 256), 256, 0, (cudaStream_t)s>>>(h, W, bias, y, B, C, P, O);
  count_launch();
  return check_launch("merge_end_kernel");
}

extern "C" int mts_merge_end_bwd(const float* dy, const float* h, const float* W, float* dh, float* dW, float* dbias,
                                 int B, int C, int P, int O, mts_stream_t s) {
  if (!dy || !h || !W || !dh || !dW || !dbias || B <= 0 || C <= 0 || P <= 0 || O <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_merge_end_bwd: bad args");
  merge_end_bwd_input_kernel<<<bw_grid((int64_t)B * C * P * O, 256), 256, 0, (cudaStream_t)s>>>(dy, W, dh, B, C, P, O);
  count_launch();
  int rc = check_launch("merge_end_bwd_input_kernel");
  if (rc) return rc;
  merge_end_bwd_weight_kernel<<<O * O * C + O, 128, 0, (cudaStream_t)s>>>(dy, h, dW, dbias, B, C, P, O);
  count_launch();
  return check_launch("merge_end_bwd_weight_kernel");
}
