// precise.cu — kernels of the evaluation parity modes ("tf32": the reference's own evaluation regime, fp32 weights with
// TF32 matmuls, tasks/base.py:19-22; "fp32": the same path with every contraction at fp32 grade via the 3xTF32 split).
// Activations stay fp32 end to end in these modes; the GEMMs are mts_gemm with ab_dtype = MTS_F32.  What lives here is
// the rest: TF32 rounding / splitting passes, the fp32 row softmax of the reprogramming scores, and an fp32 causal
// attention (plain FMA arithmetic: attention is < 1 % of the path's FLOPs and the parity modes want it exact).
#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

// ------------------------------------------------------------------------------------------
// y = nearest TF32 of x (16 bytes per thread and step; HBM bound: 8 bytes per element)
// ------------------------------------------------------------------------------------------
__global__ void round_tf32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 v = *reinterpret_cast<const float4*>(x + i);
      v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
      *reinterpret_cast<float4*>(y + i) = v;
    } else {
      for (int64_t j = i; j < n; ++j) y[j] = round_tf32(x[j]);
    }
  }
}

// x = hi + lo (+ a remainder below 2^-22 |x|), both TF32-representable: the operand pieces of the 3xTF32 contraction
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      float4 h, l;
      h.x = round_tf32(v.x); h.y = round_tf32(v.y); h.z = round_tf32(v.z); h.w = round_tf32(v.w);
      l.x = round_tf32(v.x - h.x); l.y = round_tf32(v.y - h.y); l.z = round_tf32(v.z - h.z); l.w = round_tf32(v.w - h.w);
      *reinterpret_cast<float4*>(hi + i) = h;
      *reinterpret_cast<float4*>(lo + i) = l;
    } else {
      for (int64_t j = i; j < n; ++j) {
        const float h = round_tf32(x[j]);
        hi[j] = h;
        lo[j] = round_tf32(x[j] - h);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// fp32 row softmax with scale (reprogramming scores, models/medtsllm.py:587): one warp per row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_f32_kernel(const float* __restrict__ s, float* __restrict__ p, int64_t rows, int n, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * n;
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, sr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  mx *= scale;
  float sum = 0.0f;
  for (int i = lane; i < n; i += 32) sum += expf(sr[i] * scale - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  float* pr = p + row * n;
  for (int i = lane; i < n; i += 32) pr[i] = expf(sr[i] * scale - mx) * inv;
}

// ------------------------------------------------------------------------------------------
// fp32 causal attention, eager semantics (HF:models/llama/modeling_llama.py:199-221, HF:models/gpt2/modeling_gpt2.py:
// 54-72; no padding mask, models/medtsllm.py:350), plain or shared-prefix row layout.
//
//   qkv fp32 [Lc + Bp*Ls, 3*H*HD] (q / k already rotated for Llama), out fp32 [Lc + Bp*Ls, H*HD].
//   Sequence position p of sample b lives in row  p (p < Lc)  or  Lc + b*Ls + (p - Lc).
//   grid (ceil(max(Lc, Ls) / 32), H, Bp + (Lc > 0)): block z < Bp serves the own positions [Lc, Lc + Ls) of sample z,
//   block z == Bp the shared prefix positions [0, Lc).
//   256 threads = 32 query rows x 8 lanes; per 32-key tile: S = Q K^T (each thread 4 keys), online softmax in fp32,
//   O += P V (each thread HD/8 output columns).  K is staged transposed so that both contractions read shared memory
//   conflict-free.  Roofline: FP32 FMA pipe; 4*L*L*HD FLOP per (sample, head) before causal skipping.
// ------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256)
attn_causal_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int Bp, int Lc, int Ls, int H,
                       float scale, int round_out) {
  constexpr int TQ = 32, TK = 32, CPT = HD / 8;      // output columns per thread
  // dynamic shared memory (55.9 KB at HD = 128): Qs[TQ][HD+4] (the 4 query rows a warp reads at one d fall in different
  // banks), Kt[HD][TK+4] (transposed keys; the pitch keeps float4 alignment), Vs[TK][HD], Ps[TQ][TK+1]
  extern __shared__ __align__(16) float attn_f32_smem[];
  float (*Qs)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(attn_f32_smem);
  float (*Kt)[TK + 4] = reinterpret_cast<float (*)[TK + 4]>(attn_f32_smem + TQ * (HD + 4));
  float (*Vs)[HD] = reinterpret_cast<float (*)[HD]>(attn_f32_smem + TQ * (HD + 4) + HD * (TK + 4));
  float (*Ps)[TK + 1] = reinterpret_cast<float (*)[TK + 1]>(attn_f32_smem + TQ * (HD + 4) + HD * (TK + 4) + TK * HD);
  const int z = blockIdx.z, head = blockIdx.y;
  const bool prefix = (z == Bp);
  const int q_begin = prefix ? 0 : Lc, q_end = prefix ? Lc : Lc + Ls;
  const int q0 = q_begin + blockIdx.x * TQ;
  if (q0 >= q_end) return;
  const int64_t ld = 3 * (int64_t)H * HD;
  auto row_of = [&](int pos) -> int64_t { return pos < Lc ? pos : (int64_t)Lc + (int64_t)z * Ls + (pos - Lc); };
  const int tid = threadIdx.x;
  const int r = tid >> 3, c8 = tid & 7;

  // stage the query tile
  for (int i = tid; i < TQ * (HD / 4); i += 256) {
    const int rr = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + rr < q_end) v = *reinterpret_cast<const float4*>(qkv + row_of(q0 + rr) * ld + (int64_t)head * HD + c4);
    *reinterpret_cast<float4*>(&Qs[rr][c4]) = v;
  }
  float o[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) o[j] = 0.0f;
  float m_run = -INFINITY, l_run = 0.0f;
  const int qpos = q0 + r;
  const int k_last = min(q0 + TQ - 1, q_end - 1);    // last key position any query of this tile may see

  for (int k0 = 0; k0 <= k_last; k0 += TK) {
    __syncthreads();                                  // previous tile's Kt / Vs / Ps reads are done (and Qs is staged)
    for (int i = tid; i < TK * (HD / 4); i += 256) {
      const int kk = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + kk <= k_last) {
        const float* base = qkv + row_of(k0 + kk) * ld + (int64_t)head * HD + c4;
        kv = *reinterpret_cast<const float4*>(base + (int64_t)H * HD);
        vv = *reinterpret_cast<const float4*>(base + 2 * (int64_t)H * HD);
      }
      Kt[c4][kk] = kv.x; Kt[c4 + 1][kk] = kv.y; Kt[c4 + 2][kk] = kv.z; Kt[c4 + 3][kk] = kv.w;
      *reinterpret_cast<float4*>(&Vs[kk][c4]) = vv;
    }
    __syncthreads();
    // scores of query row r against keys k0 + 4*c8 .. +3
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      const float qv = Qs[r][d];
      const float4 kv = *reinterpret_cast<const float4*>(&Kt[d][4 * c8]);
      s[0] = fmaf(qv, kv.x, s[0]); s[1] = fmaf(qv, kv.y, s[1]); s[2] = fmaf(qv, kv.z, s[2]); s[3] = fmaf(qv, kv.w, s[3]);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kpos = k0 + 4 * c8 + j;
      s[j] = (kpos <= qpos && qpos < q_end) ? s[j] * scale : -INFINITY;
      mx = fmaxf(mx, s[j]);
    }
#pragma unroll
    for (int o2 = 4; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));   // the 8 lanes of a row
    const float m_new = fmaxf(m_run, mx);
    // rows past q_end (tile padding) and fully masked tiles keep m_new = -inf: guard the exponent
    const float corr = (m_new == -INFINITY) ? 1.0f : expf(m_run - m_new);
    float psum = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float pv = (s[j] == -INFINITY) ? 0.0f : expf(s[j] - m_new);
      Ps[r][4 * c8 + j] = pv;
      psum += pv;
    }
#pragma unroll
    for (int o2 = 4; o2 > 0; o2 >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o2);
    l_run = l_run * corr + psum;
    m_run = m_new;
#pragma unroll
    for (int j = 0; j < CPT; ++j) o[j] *= corr;
    __syncwarp();                                     // a row's 8 lanes sit in one warp: Ps[r][*] is complete
#pragma unroll 4
    for (int kk = 0; kk < TK; ++kk) {
      const float pv = Ps[r][kk];
#pragma unroll
      for (int j = 0; j < CPT; j += 4) {
        // thread c8 owns columns (j/4)*32 + 4*c8 .. +3: a row's 8 lanes read 32 consecutive floats (conflict-free)
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[kk][j * 8 + 4 * c8]);
        o[j] = fmaf(pv, vv.x, o[j]); o[j + 1] = fmaf(pv, vv.y, o[j + 1]);
        o[j + 2] = fmaf(pv, vv.z, o[j + 2]); o[j + 3] = fmaf(pv, vv.w, o[j + 3]);
      }
    }
  }
  if (qpos < q_end) {
    const float inv = 1.0f / l_run;
    float* dst = out + row_of(qpos) * ((int64_t)H * HD) + (int64_t)head * HD + 4 * c8;
#pragma unroll
    for (int j = 0; j < CPT; j += 4) {
      float4 v = make_float4(o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv);
      if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      *reinterpret_cast<float4*>(dst + j * 8) = v;
    }
  }
}


// ------------------------------------------------------------------------------------------
// The same attention on the tensor cores in TF32 (mma.sync.m16n8k8) — the "tf32" mode's attention: the reference's
// evaluation regime runs Q K^T and P V as TF32 matmuls too (tasks/base.py:19-22).  q / k / v are rounded to nearest TF32
// while they are staged, P when it goes through shared memory (the m16n8k8 accumulator layout is not the A-operand
// layout: one per-warp smem round trip); softmax, running max / sum and the output stay fp32.
//   grid (ceil(max(Lc, Ls) / 64), H, Bp + (Lc > 0)), 128 threads = 4 warps x 16 query rows; 32-key tiles.
//   Shared memory pitches (HD+4 for Q / K / P rows, HD+8 for V) make every fragment load conflict-free.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD>
__global__ void __launch_bounds__(128)
attn_causal_tf32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int Bp, int Lc, int Ls, int H,
                        float scale, int round_out) {
  constexpr int TQ = 64, TK = 32, PQ = HD + 4, PVp = HD + 8, PP = TK + 4;
  extern __shared__ __align__(16) float attn_tf32_smem[];
  float* Qs = attn_tf32_smem;                 // [TQ][PQ]
  float* Ks = Qs + TQ * PQ;                   // [TK][PQ]
  float* Vs = Ks + TK * PQ;                   // [TK][PVp]
  float* Ps = Vs + TK * PVp;                  // [4 warps][16][PP]
  const int z = blockIdx.z, head = blockIdx.y;
  const bool prefix = (z == Bp);
  const int q_begin = prefix ? 0 : Lc, q_end = prefix ? Lc : Lc + Ls;
  const int q0 = q_begin + blockIdx.x * TQ;
  if (q0 >= q_end) return;
  const int64_t ld = 3 * (int64_t)H * HD;
  auto row_of = [&](int pos) -> int64_t { return pos < Lc ? pos : (int64_t)Lc + (int64_t)z * Ls + (pos - Lc); };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  for (int i = tid; i < TQ * (HD / 4); i += 128) {
    const int rr = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + rr < q_end) v = *reinterpret_cast<const float4*>(qkv + row_of(q0 + rr) * ld + (int64_t)head * HD + c4);
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    *reinterpret_cast<float4*>(Qs + rr * PQ + c4) = v;
  }
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int strip0 = q0 + warp * 16;
  const int row_a = strip0 + g, row_b = row_a + 8;
  const int k_last = min(q0 + TQ - 1, q_end - 1);
  const int strip_last = min(strip0 + 15, q_end - 1);
  const float* qa = Qs + (warp * 16 + g) * PQ;
  const float* qb = qa + 8 * PQ;
  float* pw = Ps + warp * 16 * PP;

  for (int k0 = 0; k0 <= k_last; k0 += TK) {
    __syncthreads();
    for (int i = tid; i < TK * (HD / 4); i += 128) {
      const int kk = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + kk <= k_last) {
        const float* base = qkv + row_of(k0 + kk) * ld + (int64_t)head * HD + c4;
        kv = *reinterpret_cast<const float4*>(base + (int64_t)H * HD);
        vv = *reinterpret_cast<const float4*>(base + 2 * (int64_t)H * HD);
      }
      kv.x = round_tf32(kv.x); kv.y = round_tf32(kv.y); kv.z = round_tf32(kv.z); kv.w = round_tf32(kv.w);
      vv.x = round_tf32(vv.x); vv.y = round_tf32(vv.y); vv.z = round_tf32(vv.z); vv.w = round_tf32(vv.w);
      *reinterpret_cast<float4*>(Ks + kk * PQ + c4) = kv;
      *reinterpret_cast<float4*>(Vs + kk * PVp + c4) = vv;
    }
    __syncthreads();
    if (k0 > strip_last || strip0 >= q_end) continue;          // warp-uniform: every key of this tile is masked for the strip

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < HD / 8; ++ks) {
      uint32_t a[4];
      a[0] = __float_as_uint(qa[ks * 8 + t]);
      a[1] = __float_as_uint(qb[ks * 8 + t]);
      a[2] = __float_as_uint(qa[ks * 8 + t + 4]);
      a[3] = __float_as_uint(qb[ks * 8 + t + 4]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* kr = Ks + (nt * 8 + g) * PQ + ks * 8 + t;
        mma_tf32_1688(s[nt], a, __float_as_uint(kr[0]), __float_as_uint(kr[4]));
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = k0 + nt * 8 + 2 * t + (e & 1);
        const int row = (e < 2) ? row_a : row_b;
        const float v = (col <= row && row < q_end) ? s[nt][e] * scale : -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      m_new[r] = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_new[r] == -INFINITY) ? 1.0f : expf(m_run[r] - m_new[r]);
      m_run[r] = m_new[r];
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        p[e] = (s[nt][e] == -INFINITY) ? 0.f : expf(s[nt][e] - m_new[e >> 1]);
        psum[e >> 1] += p[e];
      }
      *reinterpret_cast<float2*>(pw + g * PP + nt * 8 + 2 * t) = make_float2(round_tf32(p[0]), round_tf32(p[1]));
      *reinterpret_cast<float2*>(pw + (g + 8) * PP + nt * 8 + 2 * t) = make_float2(round_tf32(p[2]), round_tf32(p[3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 1);
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 2);
      l_run[r] = l_run[r] * corr[r] + psum[r];
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    __syncwarp();
#pragma unroll
    for (int kst = 0; kst < TK / 8; ++kst) {
      uint32_t a[4];
      a[0] = __float_as_uint(pw[g * PP + kst * 8 + t]);
      a[1] = __float_as_uint(pw[(g + 8) * PP + kst * 8 + t]);
      a[2] = __float_as_uint(pw[g * PP + kst * 8 + t + 4]);
      a[3] = __float_as_uint(pw[(g + 8) * PP + kst * 8 + t + 4]);
      const float* v0 = Vs + (kst * 8 + t) * PVp + g;
      const float* v1 = v0 + 4 * PVp;
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt)
        mma_tf32_1688(o[nt], a, __float_as_uint(v0[nt * 8]), __float_as_uint(v1[nt * 8]));
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = r ? row_b : row_a;
    if (row < q_end) {
      const float inv = 1.0f / l_run[r];
      float* dst = out + row_of(row) * ((int64_t)H * HD) + (int64_t)head * HD + 2 * t;
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) {
        float2 v = make_float2(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
        if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); }
        *reinterpret_cast<float2*>(dst + nt * 8) = v;
      }
    }
  }
}

}  // namespace mts

using namespace mts;

extern "C" int mts_round_tf32(const float* x, float* y, int64_t n, mts_stream_t stream_) {
  if (!x || !y || n < 0) return set_error(MTS_ERR_INVALID_ARG, "mts_round_tf32: null pointer / negative size");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_round_tf32: pointers must be 16-byte aligned");
  if (n == 0) return MTS_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t blocks = (n / 4 + 255) / 256 + 1;
  round_tf32_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(x, y, n);
  count_launch();
  return check_launch("round_tf32_kernel");
}

extern "C" int mts_split_tf32(const float* x, float* hi, float* lo, int64_t n, mts_stream_t stream_) {
  if (!x || !hi || !lo || n < 0) return set_error(MTS_ERR_INVALID_ARG, "mts_split_tf32: null pointer / negative size");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(hi) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_split_tf32: pointers must be 16-byte aligned");
  if (n == 0) return MTS_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t blocks = (n / 4 + 255) / 256 + 1;
  split_tf32_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(x, hi, lo, n);
  count_launch();
  return check_launch("split_tf32_kernel");
}

extern "C" int mts_softmax_rows_f32(const float* s, float* p, int64_t rows, int n, float scale, mts_stream_t stream_) {
  if (!s || !p || rows <= 0 || n <= 0) return set_error(MTS_ERR_INVALID_ARG, "mts_softmax_rows_f32: bad arguments");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t blocks = (rows + 7) / 8;
  if (blocks > 0x7fffffffL) return set_error(MTS_ERR_INVALID_ARG, "mts_softmax_rows_f32: too many rows");
  softmax_rows_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(s, p, rows, n, scale);
  count_launch();
  return check_launch("softmax_rows_f32_kernel");
}

template <int HD>
static int launch_attn_tf32(const float* qkv, float* out, int Bp, int Lc, int Ls, int H, float scale, int round_out,
                            cudaStream_t stream) {
  constexpr int kSmem = (64 * (HD + 4) + 32 * (HD + 4) + 32 * (HD + 8) + 4 * 16 * 36) * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_causal_tf32_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn_causal_tf32_kernel)", e);
    attr = true;
  }
  const int qmax = Lc > Ls ? Lc : Ls;
  dim3 grid((qmax + 63) / 64, H, Bp + (Lc > 0 ? 1 : 0));
  attn_causal_tf32_kernel<HD><<<grid, 128, kSmem, stream>>>(qkv, out, Bp, Lc, Ls, H, scale, round_out);
  count_launch();
  return check_launch("attn_causal_tf32_kernel");
}

extern "C" int mts_attn_causal_tf32(const float* qkv, float* out, int Bp, int Lc, int Ls, int H, int hd, float scale,
                                    int round_out, mts_stream_t stream_) {
  if (!qkv || !out || Bp <= 0 || Lc < 0 || Ls <= 0 || H <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_tf32: bad arguments");
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_tf32: pointers must be 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (hd == 128) return launch_attn_tf32<128>(qkv, out, Bp, Lc, Ls, H, scale, round_out, stream);
  if (hd == 64) return launch_attn_tf32<64>(qkv, out, Bp, Lc, Ls, H, scale, round_out, stream);
  return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_tf32: head dim must be 64 or 128");
}

extern "C" int mts_attn_causal_f32(const float* qkv, float* out, int Bp, int Lc, int Ls, int H, int hd, float scale,
                                   int round_out, mts_stream_t stream_) {
  if (!qkv || !out || Bp <= 0 || Lc < 0 || Ls <= 0 || H <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_f32: bad arguments");
  if (hd != 64 && hd != 128) return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_f32: head dim must be 64 or 128");
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_f32: pointers must be 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int qmax = Lc > Ls ? Lc : Ls;
  dim3 grid((qmax + 31) / 32, H, Bp + (Lc > 0 ? 1 : 0));
  auto smem_bytes = [](int HD) { return (32 * (HD + 4) + HD * 36 + 32 * HD + 32 * 33) * 4; };
  if (hd == 128) {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(attn_causal_f32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           smem_bytes(128));
      if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn_causal_f32_kernel)", e);
      attr = true;
    }
    attn_causal_f32_kernel<128><<<grid, 256, smem_bytes(128), stream>>>(qkv, out, Bp, Lc, Ls, H, scale, round_out);
  } else {
    attn_causal_f32_kernel<64><<<grid, 256, smem_bytes(64), stream>>>(qkv, out, Bp, Lc, Ls, H, scale, round_out);
  }
  count_launch();
  return check_launch("attn_causal_f32_kernel");
}
