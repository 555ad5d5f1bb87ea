// attention.cu — K8: causal self-attention of the frozen backbone with eager semantics
// (ref: HF:models/llama/modeling_llama.py:199-221 eager_attention_forward, RoPE :124-168;
//  HF:models/gpt2/modeling_gpt2.py:54-72).  No padding mask: the reference never passes one
// (models/medtsllm.py:350), so left-pad tokens are attended to like any other.
//
// Sequence lengths on this path are short (prompt + patches, L <= ~260), so attention is < 1 % of
// the FLOPs.  Per the design brief it is a shared-memory-tiled kernel with warp-level reductions:
// 64-query x 64-key tiles, Q/K/V staged in padded smem (RoPE applied while staging, in fp32),
// S = QK^T and O = PV on the warp-level tensor path (mma.sync m16n8k16 bf16 -> fp32), online
// softmax in fp32 registers with quad shuffles.  The big GEMMs are the tcgen05 kernels.
//
// Algorithmic work: 4*Bp*H*L*L*hd FLOP (dense; the causal half is skipped in practice).
#include <algorithm>
#include <cstdlib>

#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

// attention_tc.cu: the tensor-memory (tcgen05) forward for L <= 256
bool attn_tc_eligible(int L, int Lc, int hd, int Bp, int H);
int launch_attn_tc(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int L, int Lc, int H, int hd, float scale,
                   cudaStream_t stream);
bool attn_tc_bwd_eligible(int Lc, int Ls, int hd, int Bp, int H, const float* rc, const float* rs);
int launch_attn_bwd_tc(const uint16_t* qkv, const float* rc, const float* rs, const uint16_t* out_own, const uint16_t* dout_own,
                       const float* lse_own, float* delta_own, uint16_t* dqkv_own, int Bp, int Lc, int Ls, int H, int hd,
                       float scale, cudaStream_t stream);


constexpr int kAttnBlockQ = 64;
constexpr int kAttnBlockKV = 64;
constexpr int kAttnThreads = 128;

// Stage `rows` x HD bf16 from global (row stride ld elements) into padded smem [64][HD+8],
// optionally applying rotate-half RoPE with position = global row index.  Rows >= L are zeroed.
template <int HD, bool kRope>
__device__ __forceinline__ void stage_tile(__nv_bfloat16* dst, const __nv_bfloat16* __restrict__ src,
                                           int64_t ld, int row0, int L,
                                           const float* __restrict__ cosb,
                                           const float* __restrict__ sinb) {
  constexpr int kPitch = HD + 8;
  constexpr int kHalf = HD / 2;
  if (kRope) {
    constexpr int kVecPerRow = kHalf / 8;  // each item handles columns [j, j+8) and [j+HD/2, ...)
    for (int i = threadIdx.x; i < 64 * kVecPerRow; i += kAttnThreads) {
      const int r = i / kVecPerRow, j = (i - r * kVecPerRow) * 8;
      const int row = row0 + r;
      uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
      if (row < L) {
        const uint4 a = *reinterpret_cast<const uint4*>(src + (int64_t)row * ld + j);
        const uint4 b = *reinterpret_cast<const uint4*>(src + (int64_t)row * ld + kHalf + j);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        const float* cr = cosb + (int64_t)row * kHalf + j;
        const float* sr = sinb + (int64_t)row * kHalf + j;
        uint32_t lo_w[4], hi_w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float x1a = bf16_lo(aw[q]), x1b = bf16_hi(aw[q]);
          const float x2a = bf16_lo(bw[q]), x2b = bf16_hi(bw[q]);
          const float ca = cr[2 * q], cb = cr[2 * q + 1], sa = sr[2 * q], sb = sr[2 * q + 1];
          lo_w[q] = pack_bf16(x1a * ca - x2a * sa, x1b * cb - x2b * sb);
          hi_w[q] = pack_bf16(x2a * ca + x1a * sa, x2b * cb + x1b * sb);
        }
        lo = make_uint4(lo_w[0], lo_w[1], lo_w[2], lo_w[3]);
        hi = make_uint4(hi_w[0], hi_w[1], hi_w[2], hi_w[3]);
      }
      *reinterpret_cast<uint4*>(dst + r * kPitch + j) = lo;
      *reinterpret_cast<uint4*>(dst + r * kPitch + kHalf + j) = hi;
    }
  } else {
    constexpr int kVecPerRow = HD / 8;
    for (int i = threadIdx.x; i < 64 * kVecPerRow; i += kAttnThreads) {
      const int r = i / kVecPerRow, j = (i - r * kVecPerRow) * 8;
      const int row = row0 + r;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < L) v = *reinterpret_cast<const uint4*>(src + (int64_t)row * ld + j);
      *reinterpret_cast<uint4*>(dst + r * kPitch + j) = v;
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_causal_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rope_cos,
                       const float* __restrict__ rope_sin, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ lse, int L, int H, float scale_log2e) {
  constexpr int kPitch = HD + 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Ks = Qs + 64 * kPitch;
  __nv_bfloat16* Vs = Ks + 64 * kPitch;

  const int n_qblk = (L + kAttnBlockQ - 1) / kAttnBlockQ;
  const int bh = blockIdx.x / n_qblk;
  const int q0 = (blockIdx.x - bh * n_qblk) * kAttnBlockQ;
  const int b = bh / H, h = bh - b * H;
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const __nv_bfloat16* qbase = qkv + (int64_t)b * L * ld + (int64_t)h * HD;
  const __nv_bfloat16* kbase = qbase + D;
  const __nv_bfloat16* vbase = qbase + 2 * D;
  const bool rope = rope_cos != nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;

  if (rope) stage_tile<HD, true>(Qs, qbase, ld, q0, L, rope_cos, rope_sin);
  else      stage_tile<HD, false>(Qs, qbase, ld, q0, L, nullptr, nullptr);
  __syncthreads();

  // Q fragments stay in registers for the whole kernel
  uint32_t qf[HD / 16][4];
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    const uint32_t addr = smem_u32(Qs + (warp * 16 + (lane & 15)) * kPitch + ks * 16 + (lane >> 4) * 8);
    ldmatrix_x4(addr, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f; }
  float m_run[2] = {-INFINITY, -INFINITY};  // running max (log2 units), rows g and g+8
  float l_run[2] = {0.0f, 0.0f};            // per-thread partial row sums

  const int row_a = q0 + warp * 16 + g;  // this thread's two query rows
  const int row_b = row_a + 8;
  const int kv_end = min(L, q0 + kAttnBlockQ);  // causal: keys <= last query of the tile

  for (int j0 = 0; j0 < kv_end; j0 += kAttnBlockKV) {
    __syncthreads();  // previous K/V tile fully consumed
    if (rope) stage_tile<HD, true>(Ks, kbase, ld, j0, L, rope_cos, rope_sin);
    else      stage_tile<HD, false>(Ks, kbase, ld, j0, L, nullptr, nullptr);
    stage_tile<HD, false>(Vs, vbase, ld, j0, L, nullptr, nullptr);
    __syncthreads();

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f; }
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int id = lane >> 3;
        const uint32_t addr = smem_u32(Ks + (np * 16 + (id >> 1) * 8 + (lane & 7)) * kPitch +
                                       ks * 16 + (id & 1) * 8);
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4(addr, r0, r1, r2, r3);
        mma_bf16_16816(s[2 * np], qf[ks], r0, r1);
        mma_bf16_16816(s[2 * np + 1], qf[ks], r2, r3);
      }
    }

    // scale, causal mask, online softmax
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = j0 + nt * 8 + tq * 2 + (e & 1);
        const int row = (e < 2) ? row_a : row_b;
        float v = s[nt][e] * scale_log2e;
        if (col > row || col >= L) v = -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float alpha[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      m_new[r] = fmaxf(m_run[r], mx[r]);
      const float m_safe = (m_new[r] == -INFINITY) ? 0.0f : m_new[r];
      alpha[r] = exp2f(m_run[r] - m_safe);  // m_run = -inf on the first tile -> 0
      m_run[r] = m_new[r];
      m_new[r] = m_safe;
      l_run[r] *= alpha[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[nt][e] - m_new[e >> 1]);
        s[nt][e] = pv;
        l_run[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }

    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {
        const int id = lane >> 3;
        const uint32_t addr = smem_u32(Vs + (kk * 16 + (id & 1) * 8 + (lane & 7)) * kPitch +
                                       np * 16 + (id >> 1) * 8);
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4_trans(addr, r0, r1, r2, r3);
        mma_bf16_16816(o[2 * np], pa, r0, r1);
        mma_bf16_16816(o[2 * np + 1], pa, r2, r3);
      }
    }
  }

  // finalise: reduce the row sums over the quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv_a = l_run[0] > 0.0f ? 1.0f / l_run[0] : 0.0f;
  const float inv_b = l_run[1] > 0.0f ? 1.0f / l_run[1] : 0.0f;
  __nv_bfloat16* obase = out + (int64_t)b * L * D + (int64_t)h * HD;
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
    const int col = nt * 8 + tq * 2;
    if (row_a < L)
      *reinterpret_cast<uint32_t*>(obase + (int64_t)row_a * D + col) =
          pack_bf16(o[nt][0] * inv_a, o[nt][1] * inv_a);
    if (row_b < L)
      *reinterpret_cast<uint32_t*>(obase + (int64_t)row_b * D + col) =
          pack_bf16(o[nt][2] * inv_b, o[nt][3] * inv_b);
  }
  if (lse && tq == 0) {
    // natural-log LSE of the scaled scores: ln(sum exp(s*scale)) = (m + log2(l)) * ln 2
    if (row_a < L) lse[(int64_t)bh * L + row_a] = (m_run[0] + log2f(l_run[0])) * 0.6931471805599453f;
    if (row_b < L) lse[(int64_t)bh * L + row_b] = (m_run[1] + log2f(l_run[1])) * 0.6931471805599453f;
  }
}


// ---------------------------------------------------------------------------------------------
// RoPE in place on the q and k sections of a [Bp*L, 3*H*hd] bf16 buffer (rotate-half, position =
// row index inside the sample).  16 bytes per access; bytes: 8*M*D read + written.
// ---------------------------------------------------------------------------------------------
__global__ void rope_qk_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ cosb,
                               const float* __restrict__ sinb, int64_t rows, int L, int Lc, int H, int hd) {
  pdl_wait();
  pdl_trigger();
  const int half = hd >> 1;
  const int vec_per_head = half >> 3;                 // 8 column pairs per item
  const int64_t per_row = (int64_t)2 * H * vec_per_head;  // q heads then k heads
  const int64_t total = rows * per_row;
  const int64_t D = (int64_t)H * hd;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / per_row;
    const int rem = (int)(idx - row * per_row);
    const int head = rem / vec_per_head;              // 0..2H-1 (>= H: key heads)
    const int j = (rem - head * vec_per_head) * 8;
    const int pos = row < Lc ? (int)row : Lc + (int)((row - Lc) % L);   // Lc: shared-prefix rows (0 = plain layout)
    __nv_bfloat16* base = qkv + row * 3 * D + (int64_t)head * hd;  // k section follows q contiguously
    const uint4 a = *reinterpret_cast<const uint4*>(base + j);
    const uint4 b = *reinterpret_cast<const uint4*>(base + half + j);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    const float* cr = cosb + (int64_t)pos * half + j;
    const float* sr = sinb + (int64_t)pos * half + j;
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x1a = bf16_lo(aw[q]), x1b = bf16_hi(aw[q]);
      const float x2a = bf16_lo(bw[q]), x2b = bf16_hi(bw[q]);
      const float ca = cr[2 * q], cb = cr[2 * q + 1], sa = sr[2 * q], sb = sr[2 * q + 1];
      lo[q] = pack_bf16(x1a * ca - x2a * sa, x1b * cb - x2b * sb);
      hi[q] = pack_bf16(x2a * ca + x1a * sa, x2b * cb + x1b * sb);
    }
    *reinterpret_cast<uint4*>(base + j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(base + half + j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------
// Forward, short sequences (the K and V a CTA needs fit in shared memory, L <= ~350 at hd = 128): K/V
// staged ONCE with cp.async (q/k already rotated by the qkv GEMM epilogue / rope_qk_kernel), 8 warps pull
// 16-query strips from a shared counter, heaviest (latest) strips first.  Key tiles beyond a strip's causal
// limit are skipped at 16-key granularity.
//
// Row layouts.  Plain (Lc = 0): sample b owns rows b*L + [0, L); one CTA per (b, h).
// Shared prefix (Lc > 0): rows [0, Lc) hold ONE copy of the prompt prefix every sample shares (positions
// 0..Lc-1), rows Lc + b*Ls + t the own token t of sample b at position Lc + t (Ls = L - Lc).  A CTA then
// serves `spc` consecutive samples of one head: the prefix K/V are staged once for all of them, followed by
// the samples' own K/V (shared-memory row of key position p of sample i: p if p < Lc, else p + i*Ls — the
// same formula as the global row, which is why the staging loop is one contiguous copy), and the strips of
// all spc samples feed the 8 warps.  The first H CTAs of a shared-prefix launch compute the prefix itself
// (an ordinary causal sequence of Lc positions).
// ---------------------------------------------------------------------------------------------
constexpr int kSeqThreads = 256;

template <int HD>
__global__ void __launch_bounds__(kSeqThreads)
attn_causal_fwd_seq_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                           float* __restrict__ lse, int Bp, int L_all, int Lc_all, int spc, int rows_alloc, int H,
                           float scale_log2e, uint32_t drop_thresh, float drop_scale, uint64_t drop_seed) {
  // drop_thresh != 0 (plain layout only): dropout on the attention probabilities after the softmax
  // (HF:models/gpt2/modeling_gpt2.py:67-68, HF:models/llama/modeling_llama.py:217): the normaliser l and the saved
  // log-sum-exp use the unmasked P, the P V contraction the masked one; element (b, h, q, k) has index
  // ((b*H + h)*L + q)*L + k in the counter-based mask.
  pdl_wait();
  pdl_trigger();
  constexpr int kPitch = HD + 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Vs = Ks + (size_t)rows_alloc * kPitch;
  __nv_bfloat16* Qw = Vs + (size_t)rows_alloc * kPitch;           // [8 warps][16][kPitch]
  int* counter = reinterpret_cast<int*>(Qw + 8 * 16 * kPitch);

  // job of this CTA: (L keys per sample, of which Lc shared; samples b0 .. b0+nb-1; head h)
  int L = L_all, Lc = Lc_all, b0, nb, h;
  float* lse_job = lse;
  if (Lc_all > 0 && (int)blockIdx.x < H) {          // the prefix as one sequence
    h = blockIdx.x; L = Lc_all; Lc = 0; b0 = 0; nb = 1;
  } else {
    const int j = Lc_all > 0 ? blockIdx.x - H : blockIdx.x;
    const int grp = j / H;
    h = j - grp * H;
    b0 = grp * spc;
    nb = min(spc, Bp - b0);
    if (lse && Lc_all > 0) lse_job = lse + (int64_t)H * Lc_all;
  }
  const int Ls = L - Lc;
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  // global row of (sample b, position p): p (p < Lc) or p + b*Ls; shared-memory row of (i = b - b0, p): p or p + i*Ls
  const __nv_bfloat16* gbase = qkv + (int64_t)h * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;

  constexpr int kVec = HD / 8;
  const int rows_used = Lc + nb * Ls;
  for (int i = threadIdx.x; i < rows_alloc * kVec; i += kSeqThreads) {
    const int r = i / kVec, c = (i - r * kVec) * 8;
    if (r < rows_used) {
      const __nv_bfloat16* src = gbase + ((int64_t)r + (r < Lc ? 0 : (int64_t)b0 * Ls)) * ld + c;
      cp_async16(smem_u32(Ks + r * kPitch + c), src + D);
      cp_async16(smem_u32(Vs + r * kPitch + c), src + 2 * D);
    } else {
      *reinterpret_cast<uint4*>(Ks + r * kPitch + c) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Vs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_commit();
  if (threadIdx.x == 0) *counter = 0;
  cp_async_wait<0>();
  __syncthreads();

  const int n_strips = (Ls + 15) >> 4;              // per sample
  __nv_bfloat16* Qs = Qw + warp * 16 * kPitch;
  while (true) {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(counter, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= n_strips * nb) break;
    const int si = ticket % nb;                                   // sample within the CTA
    const int q0 = Lc + (n_strips - 1 - ticket / nb) * 16;        // heaviest strips first (positions, >= Lc)
    const int soff = si * Ls;                                     // shared-memory row offset of own keys
    const int64_t grow = (int64_t)(b0 + si) * Ls;                 // global row offset of own rows
    // stage this warp's 16 query rows
    for (int i = lane; i < 16 * kVec; i += 32) {
      const int r = i / kVec, c = (i - r * kVec) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (q0 + r < L) v = *reinterpret_cast<const uint4*>(gbase + (grow + q0 + r) * ld + c);
      *reinterpret_cast<uint4*>(Qs + r * kPitch + c) = v;
    }
    __syncwarp();
    uint32_t qf[HD / 16][4];
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      ldmatrix_x4(smem_u32(Qs + (lane & 15) * kPitch + ks * 16 + (lane >> 4) * 8), qf[ks][0], qf[ks][1],
                  qf[ks][2], qf[ks][3]);

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};
    const int row_a = q0 + g, row_b = row_a + 8;
    const int last_row = min(q0 + 15, L - 1);

    for (int j0 = 0; j0 <= last_row; j0 += 64) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f; }
      // k-step outer, key group inner: the 8 MMAs of a k-step hit 8 different accumulators (no dependent chain) and the
      // 4 K-fragment loads of a k-step are issued back to back ahead of them
      const int n_grp = min(4, (last_row - j0) / 16 + 1);                   // warp-uniform causal skip
      uint32_t krow_addr[4];
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int id = lane >> 3;
        const int kp = j0 + np * 16 + (id >> 1) * 8 + (lane & 7);          // key position of this lane's row
        krow_addr[np] = smem_u32(Ks + (kp + (kp < Lc ? 0 : soff)) * kPitch + (id & 1) * 8);
      }
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t kf[4][4];
#pragma unroll
        for (int np = 0; np < 4; ++np)
          if (np < n_grp) ldmatrix_x4(krow_addr[np] + ks * 32, kf[np][0], kf[np][1], kf[np][2], kf[np][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          if (np < n_grp) {
            mma_bf16_16816(s[2 * np], qf[ks], kf[np][0], kf[np][1]);
            mma_bf16_16816(s[2 * np + 1], qf[ks], kf[np][2], kf[np][3]);
          }
        }
      }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = j0 + nt * 8 + tq * 2 + (e & 1);
          const int row = (e < 2) ? row_a : row_b;
          float v = s[nt][e] * scale_log2e;
          if (col > row || col >= L) v = -INFINITY;
          s[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
      float alpha[2], m_new[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        m_new[r] = fmaxf(m_run[r], mx[r]);
        const float m_safe = (m_new[r] == -INFINITY) ? 0.0f : m_new[r];
        alpha[r] = exp2f(m_run[r] - m_safe);
        m_run[r] = m_new[r];
        m_new[r] = m_safe;
        l_run[r] *= alpha[r];
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f(s[nt][e] - m_new[e >> 1]);
          s[nt][e] = pv;
          l_run[e >> 1] += pv;
        }
      }
      if (drop_thresh != 0) {
        const uint64_t base = ((uint64_t)(b0 + si) * H + h) * (uint64_t)L * (uint64_t)L;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = j0 + nt * 8 + tq * 2 + (e & 1);
            const int row = (e < 2) ? row_a : row_b;
            s[nt][e] = dropout_keep(drop_seed, base + (uint64_t)row * L + col, drop_thresh) ? s[nt][e] * drop_scale : 0.0f;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
        o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (j0 + kk * 16 > last_row) continue;   // P is identically zero there
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const int id = lane >> 3;
        const int kp = j0 + kk * 16 + (id & 1) * 8 + (lane & 7);
        const uint32_t vaddr = smem_u32(Vs + (kp + (kp < Lc ? 0 : soff)) * kPitch + (id >> 1) * 8);
        uint32_t vf[2][4];                       // double-buffered V fragments: load np+1 ahead of the MMAs of np
        ldmatrix_x4_trans(vaddr, vf[0][0], vf[0][1], vf[0][2], vf[0][3]);
#pragma unroll
        for (int np = 0; np < HD / 16; ++np) {
          if (np + 1 < HD / 16)
            ldmatrix_x4_trans(vaddr + (np + 1) * 32, vf[(np + 1) & 1][0], vf[(np + 1) & 1][1], vf[(np + 1) & 1][2],
                              vf[(np + 1) & 1][3]);
          mma_bf16_16816(o[2 * np], pa, vf[np & 1][0], vf[np & 1][1]);
          mma_bf16_16816(o[2 * np + 1], pa, vf[np & 1][2], vf[np & 1][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv_a = l_run[0] > 0.0f ? 1.0f / l_run[0] : 0.0f;
    const float inv_b = l_run[1] > 0.0f ? 1.0f / l_run[1] : 0.0f;
    __nv_bfloat16* obase = out + grow * D + (int64_t)h * HD;
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
      const int col = nt * 8 + tq * 2;
      if (row_a < L)
        *reinterpret_cast<uint32_t*>(obase + (int64_t)row_a * D + col) = pack_bf16(o[nt][0] * inv_a, o[nt][1] * inv_a);
      if (row_b < L)
        *reinterpret_cast<uint32_t*>(obase + (int64_t)row_b * D + col) = pack_bf16(o[nt][2] * inv_b, o[nt][3] * inv_b);
    }
    if (lse_job && tq == 0) {   // lse is [samples, H, Ls], indexed by the token's place among the sample's own tokens
      float* lrow = lse_job + ((int64_t)(b0 + si) * H + h) * Ls - Lc;
      if (row_a < L) lrow[row_a] = (m_run[0] + log2f(l_run[0])) * 0.6931471805599453f;
      if (row_b < L) lrow[row_b] = (m_run[1] + log2f(l_run[1])) * 0.6931471805599453f;
    }
    __syncwarp();   // Qs is re-staged by the next strip
  }
}

// rows of K (and of V) a CTA keeps: the positions of spc samples sharing Lc of their L; the 64-key tiles of the
// last sample may overrun its L positions into a zero tail
static int seq_rows_alloc(int L, int Lc, int spc) { return ((L + 63) & ~63) + (spc - 1) * (L - Lc); }

template <int HD>
static size_t seq_smem_bytes(int L, int Lc = 0, int spc = 1) {
  return (size_t)(2 * seq_rows_alloc(L, Lc, spc) + 8 * 16) * (HD + 8) * 2 + 16;
}

// Sequence-resident forward over Bp samples of L positions whose first Lc are a shared prefix stored once
// (Lc > 0: the launch also computes the prefix rows themselves; lse = [H, Lc] followed by [Bp, H, Ls]).
template <int HD>
static int launch_attn_seq(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int L, int Lc, int H,
                           float scale, cudaStream_t stream, uint32_t drop_thresh = 0, float drop_scale = 1.0f,
                           uint64_t drop_seed = 0) {
  auto ks = attn_causal_fwd_seq_kernel<HD>;
  static bool seq_attr_done = false;
  if (!seq_attr_done) {
    cudaError_t e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn seq)", e);
    seq_attr_done = true;
  }
  // samples per CTA: enough 16-query strips to occupy the 8 warps, as far as shared memory allows
  int spc = 1;
  if (Lc > 0) {
    const int n_strips = (L - Lc + 15) / 16;
    const int want = std::min(Bp, (8 + n_strips - 1) / n_strips);
    while (spc < want && seq_smem_bytes<HD>(L, Lc, spc + 1) <= 220 * 1024) ++spc;
  }
  const int groups = (Bp + spc - 1) / spc;
  const int grid = (groups + (Lc > 0 ? 1 : 0)) * H;
  const int rows_alloc = std::max(seq_rows_alloc(L, Lc, spc), Lc > 0 ? seq_rows_alloc(Lc, 0, 1) : 0);
  const size_t smem = (size_t)(2 * rows_alloc + 8 * 16) * (HD + 8) * 2 + 16;
  LAUNCH_PDL(ks, grid, kSeqThreads, smem, stream,
      
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), lse, Bp, L, Lc, spc,
      rows_alloc, H, scale * 1.4426950408889634f, drop_thresh, drop_scale, drop_seed);
  count_launch();
  return check_launch("attn_causal_fwd_seq_kernel");
}

template <int HD>
static int launch_attn(const uint16_t* qkv, const float* rc, const float* rs, uint16_t* out,
                       float* lse, int Bp, int L, int H, float scale, cudaStream_t stream) {
  if (rc == nullptr && attn_tc_eligible(L, 0, HD, Bp, H)) return launch_attn_tc(qkv, out, lse, Bp, L, 0, H, HD, scale, stream);
  if (rc == nullptr && seq_smem_bytes<HD>(L) <= 220 * 1024)
    return launch_attn_seq<HD>(qkv, out, lse, Bp, L, 0, H, scale, stream);
  constexpr int kSmem = 3 * 64 * (HD + 8) * 2;
  auto kern = attn_causal_fwd_kernel<HD>;
  static bool attr_done = false;
  if (!attr_done && kSmem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn)", e);
    attr_done = true;
  }
  const int64_t grid_l = (int64_t)((L + kAttnBlockQ - 1) / kAttnBlockQ) * Bp * H;
  if (grid_l > 0x7fffffffLL) return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal: grid too large");
  const int grid = (int)grid_l;
  kern<<<grid, kAttnThreads, kSmem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs,
                                               reinterpret_cast<__nv_bfloat16*>(out), lse, L, H,
                                               scale * 1.4426950408889634f);
  count_launch();
  return check_launch("attn_causal_fwd_kernel");
}


// =============================================================================================
// Backward (training path: gradients flow through the frozen backbone to the adapters in front of
// it).  Two kernels, both recomputing S = QK^T from the saved qkv and the forward's log-sum-exp, so
// that neither needs atomics (deterministic):
//   attn_bwd_dq_kernel   one CTA per (b, h, 64 queries): dQ = sum_k dS K
//   attn_bwd_dkv_kernel  one CTA per (b, h, 64 keys):    dK = sum_q dS^T Q,  dV = sum_q P^T dO
// with P = exp(S*scale - lse), dS = P * (dP - delta) * scale, dP = dO V^T, delta = rowsum(dO * O).
// RoPE: q/k are rotated while staging exactly as in the forward; dQ/dK are rotated back on store.
// =============================================================================================

__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                      float* __restrict__ delta, int L, int H, int hd, int64_t total) {
  pdl_wait();
  pdl_trigger();
  // one warp per (b, l, h); delta laid out [b, h, l]
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= total) return;
  const int lane = threadIdx.x & 31;
  const int h = (int)(wid % H);
  const int64_t bl = wid / H;
  const int l = (int)(bl % L);
  const int64_t b = bl / L;
  const __nv_bfloat16* op = o + (bl * H + h) * hd;
  const __nv_bfloat16* dp = dout + (bl * H + h) * hd;
  float acc = 0.f;
  for (int i = lane * 2; i < hd; i += 64) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(op + i);
    const uint32_t d = *reinterpret_cast<const uint32_t*>(dp + i);
    acc += bf16_lo(a) * bf16_lo(d) + bf16_hi(a) * bf16_hi(d);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) delta[(b * H + h) * L + l] = acc;
}

// C(16 x 64) = A(16 x HD, rows [arow0, +16) of As) * B^T, B = 64 rows of Bs (both [row][HD+8] bf16)
// Rows of Bs at or beyond `thr` (relative to Bs) lie `shift` rows further down: the shared-prefix layout keeps the
// own keys of sample i of a CTA i*Ls rows behind the prefix keys (thr = INT_MAX: plain contiguous rows).
// kBatch: issue the 4 B-fragment loads of a k-step ahead of its 8 MMAs (16 more live registers: the dK/dV loops, which
// carry two HD-wide accumulators, stay on the interleaved order).
template <int HD, bool kBatch = true>
__device__ __forceinline__ void mma_a_bt(float (&c)[8][4], const __nv_bfloat16* As, int arow0,
                                         const __nv_bfloat16* Bs, int lane, int g_lo = 0, int g_hi = 4,
                                         int thr = 0x7fffffff, int shift = 0) {
  constexpr int kPitch = HD + 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    uint32_t a[4];
    ldmatrix_x4(smem_u32(As + (arow0 + (lane & 15)) * kPitch + ks * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
    if constexpr (kBatch) {
      uint32_t bf[4][4];                         // the B fragments of this k-step first, then the 8 independent MMAs
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np < g_lo || np >= g_hi) continue;   // warp-uniform: 16-column groups outside the causal range
        const int id = lane >> 3;
        int brow = np * 16 + (id >> 1) * 8 + (lane & 7);
        brow += brow >= thr ? shift : 0;
        ldmatrix_x4(smem_u32(Bs + brow * kPitch + ks * 16 + (id & 1) * 8), bf[np][0], bf[np][1], bf[np][2], bf[np][3]);
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np < g_lo || np >= g_hi) continue;
        mma_bf16_16816(c[2 * np], a, bf[np][0], bf[np][1]);
        mma_bf16_16816(c[2 * np + 1], a, bf[np][2], bf[np][3]);
      }
    } else {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np < g_lo || np >= g_hi) continue;
        const int id = lane >> 3;
        int brow = np * 16 + (id >> 1) * 8 + (lane & 7);
        brow += brow >= thr ? shift : 0;
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4(smem_u32(Bs + brow * kPitch + ks * 16 + (id & 1) * 8), r0, r1, r2, r3);
        mma_bf16_16816(c[2 * np], a, r0, r1);
        mma_bf16_16816(c[2 * np + 1], a, r2, r3);
      }
    }
  }
}

// acc(16 x HD) += P(16 x 64, fp32 C-fragments) * B, B = 64 rows of Bs ([row][HD+8] bf16)
template <int HD>
__device__ __forceinline__ void mma_p_b(float (&acc)[HD / 8][4], const float (&pm)[8][4],
                                        const __nv_bfloat16* Bs, int lane, int g_lo = 0, int g_hi = 4,
                                        int thr = 0x7fffffff, int shift = 0) {
  constexpr int kPitch = HD + 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (kk < g_lo || kk >= g_hi) continue;     // P is identically zero in those 16-wide groups
    uint32_t pa[4];
    pa[0] = pack_bf16(pm[2 * kk][0], pm[2 * kk][1]);
    pa[1] = pack_bf16(pm[2 * kk][2], pm[2 * kk][3]);
    pa[2] = pack_bf16(pm[2 * kk + 1][0], pm[2 * kk + 1][1]);
    pa[3] = pack_bf16(pm[2 * kk + 1][2], pm[2 * kk + 1][3]);
    const int id = lane >> 3;
    int brow = kk * 16 + (id & 1) * 8 + (lane & 7);
    brow += brow >= thr ? shift : 0;
    const uint32_t baddr = smem_u32(Bs + brow * kPitch + (id >> 1) * 8);
    uint32_t bf[2][4];                           // double-buffered: fragment np+1 is in flight during the MMAs of np
    ldmatrix_x4_trans(baddr, bf[0][0], bf[0][1], bf[0][2], bf[0][3]);
#pragma unroll
    for (int np = 0; np < HD / 16; ++np) {
      if (np + 1 < HD / 16)
        ldmatrix_x4_trans(baddr + (np + 1) * 32, bf[(np + 1) & 1][0], bf[(np + 1) & 1][1], bf[(np + 1) & 1][2],
                          bf[(np + 1) & 1][3]);
      mma_bf16_16816(acc[2 * np], pa, bf[np & 1][0], bf[np & 1][1]);
      mma_bf16_16816(acc[2 * np + 1], pa, bf[np & 1][2], bf[np & 1][3]);
    }
  }
}

// Rotate a gradient w.r.t. rotated q/k back to the un-rotated projection output and store it (bf16).
template <int HD>
__device__ __forceinline__ void store_grad_rows(float (&acc)[HD / 8][4], __nv_bfloat16* dst, int64_t ld,
                                                int row_a, int row_b, int L, int tq,
                                                const float* __restrict__ cosb, const float* __restrict__ sinb) {
  constexpr int kHalfTiles = HD / 16;
  if (cosb != nullptr) {
#pragma unroll
    for (int nt = 0; nt < kHalfTiles; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = (e < 2) ? row_a : row_b;
        if (row < L) {
          const int j = nt * 8 + tq * 2 + (e & 1);
          const float c = cosb[(int64_t)row * (HD / 2) + j], s = sinb[(int64_t)row * (HD / 2) + j];
          const float lo = acc[nt][e], hi = acc[nt + kHalfTiles][e];
          acc[nt][e] = lo * c + hi * s;
          acc[nt + kHalfTiles][e] = hi * c - lo * s;
        }
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
    const int col = nt * 8 + tq * 2;
    if (row_a < L) *reinterpret_cast<uint32_t*>(dst + (int64_t)row_a * ld + col) = pack_bf16(acc[nt][0], acc[nt][1]);
    if (row_b < L) *reinterpret_cast<uint32_t*>(dst + (int64_t)row_b * ld + col) = pack_bf16(acc[nt][2], acc[nt][3]);
  }
}

template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rope_cos,
                   const float* __restrict__ rope_sin, const __nv_bfloat16* __restrict__ dout,
                   const float* __restrict__ lse, const float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, int L, int H, float scale, int pre_roped) {
  constexpr int kPitch = HD + 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* dOs = Qs + 64 * kPitch;
  __nv_bfloat16* Ks = dOs + 64 * kPitch;
  __nv_bfloat16* Vs = Ks + 64 * kPitch;

  const int n_qblk = (L + kAttnBlockQ - 1) / kAttnBlockQ;
  const int bh = blockIdx.x / n_qblk;
  const int q0 = (blockIdx.x - bh * n_qblk) * kAttnBlockQ;
  const int b = bh / H, h = bh - b * H;
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const __nv_bfloat16* qbase = qkv + (int64_t)b * L * ld + (int64_t)h * HD;
  const __nv_bfloat16* kbase = qbase + D;
  const __nv_bfloat16* vbase = qbase + 2 * D;
  const __nv_bfloat16* dobase = dout + (int64_t)b * L * D + (int64_t)h * HD;
  const bool rope = rope_cos != nullptr && !pre_roped;   // rotate while staging?
  const bool unrope = rope_cos != nullptr;               // rotate the gradients back on store
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const float scale_log2e = scale * 1.4426950408889634f;

  if (rope) stage_tile<HD, true>(Qs, qbase, ld, q0, L, rope_cos, rope_sin);
  else      stage_tile<HD, false>(Qs, qbase, ld, q0, L, nullptr, nullptr);
  stage_tile<HD, false>(dOs, dobase, D, q0, L, nullptr, nullptr);

  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  const float lse_a = row_a < L ? lse[(int64_t)bh * L + row_a] * 1.4426950408889634f : INFINITY;
  const float lse_b = row_b < L ? lse[(int64_t)bh * L + row_b] * 1.4426950408889634f : INFINITY;
  const float del_a = row_a < L ? delta[(int64_t)bh * L + row_a] : 0.f;
  const float del_b = row_b < L ? delta[(int64_t)bh * L + row_b] : 0.f;

  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

  const int kv_end = min(L, q0 + kAttnBlockQ);
  for (int j0 = 0; j0 < kv_end; j0 += kAttnBlockKV) {
    __syncthreads();
    if (rope) stage_tile<HD, true>(Ks, kbase, ld, j0, L, rope_cos, rope_sin);
    else      stage_tile<HD, false>(Ks, kbase, ld, j0, L, nullptr, nullptr);
    stage_tile<HD, false>(Vs, vbase, ld, j0, L, nullptr, nullptr);
    __syncthreads();
    float s[8][4], dp[8][4];
    mma_a_bt<HD>(s, Qs, warp * 16, Ks, lane);
    mma_a_bt<HD>(dp, dOs, warp * 16, Vs, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = j0 + nt * 8 + tq * 2 + (e & 1);
        const int row = (e < 2) ? row_a : row_b;
        const float pv = (col > row || col >= L) ? 0.f : exp2f(s[nt][e] * scale_log2e - ((e < 2) ? lse_a : lse_b));
        s[nt][e] = pv * (dp[nt][e] - ((e < 2) ? del_a : del_b)) * scale;   // dS
      }
    }
    mma_p_b<HD>(dq, s, Ks, lane);
  }
  store_grad_rows<HD>(dq, dqkv + (int64_t)b * L * ld + (int64_t)h * HD, ld, row_a, row_b, L, tq,
                      unrope ? rope_cos : nullptr, rope_sin);
}

template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rope_cos,
                    const float* __restrict__ rope_sin, const __nv_bfloat16* __restrict__ dout,
                    const float* __restrict__ lse, const float* __restrict__ delta,
                    __nv_bfloat16* __restrict__ dqkv, int L, int H, float scale, int pre_roped) {
  constexpr int kPitch = HD + 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Vs = Ks + 64 * kPitch;
  __nv_bfloat16* Qs = Vs + 64 * kPitch;
  __nv_bfloat16* dOs = Qs + 64 * kPitch;
  float* lse_s = reinterpret_cast<float*>(dOs + 64 * kPitch);
  float* del_s = lse_s + 64;

  const int n_kblk = (L + kAttnBlockKV - 1) / kAttnBlockKV;
  const int bh = blockIdx.x / n_kblk;
  const int j0 = (blockIdx.x - bh * n_kblk) * kAttnBlockKV;
  const int b = bh / H, h = bh - b * H;
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const __nv_bfloat16* qbase = qkv + (int64_t)b * L * ld + (int64_t)h * HD;
  const __nv_bfloat16* kbase = qbase + D;
  const __nv_bfloat16* vbase = qbase + 2 * D;
  const __nv_bfloat16* dobase = dout + (int64_t)b * L * D + (int64_t)h * HD;
  const bool rope = rope_cos != nullptr && !pre_roped;   // rotate while staging?
  const bool unrope = rope_cos != nullptr;               // rotate the gradients back on store
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const float scale_log2e = scale * 1.4426950408889634f;

  if (rope) stage_tile<HD, true>(Ks, kbase, ld, j0, L, rope_cos, rope_sin);
  else      stage_tile<HD, false>(Ks, kbase, ld, j0, L, nullptr, nullptr);
  stage_tile<HD, false>(Vs, vbase, ld, j0, L, nullptr, nullptr);

  const int key_a = j0 + warp * 16 + g, key_b = key_a + 8;   // this thread's two key rows
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }

  // causal: only query blocks at or after this key block contribute
  for (int i0 = j0; i0 < L; i0 += kAttnBlockQ) {
    __syncthreads();
    if (rope) stage_tile<HD, true>(Qs, qbase, ld, i0, L, rope_cos, rope_sin);
    else      stage_tile<HD, false>(Qs, qbase, ld, i0, L, nullptr, nullptr);
    stage_tile<HD, false>(dOs, dobase, D, i0, L, nullptr, nullptr);
    if (threadIdx.x < 64) {
      const int r = i0 + threadIdx.x;
      lse_s[threadIdx.x] = r < L ? lse[(int64_t)bh * L + r] * 1.4426950408889634f : INFINITY;
      del_s[threadIdx.x] = r < L ? delta[(int64_t)bh * L + r] : 0.f;
    }
    __syncthreads();
    float st[8][4], dpt[8][4];              // rows = keys, cols = queries
    mma_a_bt<HD>(st, Ks, warp * 16, Qs, lane);
    mma_a_bt<HD>(dpt, Vs, warp * 16, dOs, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = nt * 8 + tq * 2 + (e & 1);       // query index inside the tile
        const int qrow = i0 + qi;
        const int key = (e < 2) ? key_a : key_b;
        const float pv = (key > qrow || key >= L || qrow >= L) ? 0.f
                                                               : exp2f(st[nt][e] * scale_log2e - lse_s[qi]);
        st[nt][e] = pv;                                              // P^T
        dpt[nt][e] = pv * (dpt[nt][e] - del_s[qi]) * scale;          // dS^T
      }
    }
    mma_p_b<HD>(dv, st, dOs, lane);
    mma_p_b<HD>(dk, dpt, Qs, lane);
  }
  __nv_bfloat16* dbase = dqkv + (int64_t)b * L * ld + (int64_t)h * HD;
  store_grad_rows<HD>(dk, dbase + D, ld, key_a, key_b, L, tq, unrope ? rope_cos : nullptr, rope_sin);
  store_grad_rows<HD>(dv, dbase + 2 * D, ld, key_a, key_b, L, tq, nullptr, nullptr);
}


// ---------------------------------------------------------------------------------------------
// Backward, short sequences: same idea as attn_causal_fwd_seq_kernel.  One CTA per (b, h) keeps the
// operands every strip needs resident in shared memory (dQ: K and V; dK/dV: Q, dO, lse, delta), staged
// once with cp.async from the already rotated qkv, and 8 warps pull 16-row strips from a shared counter,
// heaviest first.  Key / query groups outside a strip's causal range are skipped at 16-row granularity.
// ---------------------------------------------------------------------------------------------
template <int HD>
__device__ __forceinline__ void stage_rows_async(__nv_bfloat16* dst, const __nv_bfloat16* __restrict__ src, int64_t ld,
                                                 int L, int Lp, int tid, int nthreads) {
  constexpr int kPitch = HD + 8;
  constexpr int kVec = HD / 8;
  for (int i = tid; i < Lp * kVec; i += nthreads) {
    const int r = i / kVec, c = (i - r * kVec) * 8;
    if (r < L) cp_async16(smem_u32(dst + r * kPitch + c), src + (int64_t)r * ld + c);
    else       *reinterpret_cast<uint4*>(dst + r * kPitch + c) = make_uint4(0, 0, 0, 0);
  }
}
template <int HD>
__device__ __forceinline__ void stage_strip(__nv_bfloat16* dst, const __nv_bfloat16* __restrict__ src, int64_t ld,
                                            int row0, int L, int lane) {
  constexpr int kPitch = HD + 8;
  constexpr int kVec = HD / 8;
  for (int i = lane; i < 16 * kVec; i += 32) {
    const int r = i / kVec, c = (i - r * kVec) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row0 + r < L) v = *reinterpret_cast<const uint4*>(src + (int64_t)(row0 + r) * ld + c);
    *reinterpret_cast<uint4*>(dst + r * kPitch + c) = v;
  }
}

template <int HD>
__global__ void __launch_bounds__(kSeqThreads)
attn_bwd_dq_seq_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rope_cos,
                       const float* __restrict__ rope_sin, const __nv_bfloat16* __restrict__ dout,
                       const float* __restrict__ lse, const float* __restrict__ delta,
                       __nv_bfloat16* __restrict__ dqkv, int Bp, int L, int Lc, int spc, int rows_alloc, int H,
                       float scale, uint32_t drop_thresh, float drop_scale, uint64_t drop_seed) {
  pdl_wait();
  pdl_trigger();
  // drop_thresh != 0 (plain layout): the forward dropped attention probabilities (see attn_causal_fwd_seq_kernel):
  // dP = mask * (dO V^T), dS = P * (dP - delta) with delta = rowsum(dO * O) unchanged.
  // Lc > 0: shared-prefix layout as in attn_causal_fwd_seq_kernel (a CTA serves spc samples of one head; the prefix
  // K/V are staged once).  qkv holds every row (prefix rows once, then Ls = L - Lc own rows per sample); dout / lse /
  // delta / dqkv hold the samples' own rows only ([Bp*Ls, ...]): the prefix has no trainable ancestor, so no gradient
  // is produced for it.
  constexpr int kPitch = HD + 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Vs = Ks + (size_t)rows_alloc * kPitch;
  __nv_bfloat16* Qw = Vs + (size_t)rows_alloc * kPitch;    // [8 warps][16][kPitch]
  __nv_bfloat16* dOw = Qw + 8 * 16 * kPitch;               // [8 warps][16][kPitch]
  int* counter = reinterpret_cast<int*>(dOw + 8 * 16 * kPitch);

  const int grp = blockIdx.x / H, h = blockIdx.x - grp * H;
  const int b0 = grp * spc, nb = min(spc, Bp - b0);
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const int Ls = L - Lc;
  const __nv_bfloat16* gbase = qkv + (int64_t)h * HD;      // row of (sample b, position p): p (p < Lc) or p + b*Ls
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const float scale_log2e = scale * 1.4426950408889634f;

  {
    constexpr int kVec = HD / 8;
    const int rows_used = Lc + nb * Ls;
    for (int i = threadIdx.x; i < rows_alloc * kVec; i += kSeqThreads) {
      const int r = i / kVec, c = (i - r * kVec) * 8;
      if (r < rows_used) {
        const __nv_bfloat16* src = gbase + ((int64_t)r + (r < Lc ? 0 : (int64_t)b0 * Ls)) * ld + c;
        cp_async16(smem_u32(Ks + r * kPitch + c), src + D);
        cp_async16(smem_u32(Vs + r * kPitch + c), src + 2 * D);
      } else {
        *reinterpret_cast<uint4*>(Ks + r * kPitch + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Vs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  cp_async_commit();
  if (threadIdx.x == 0) *counter = 0;
  cp_async_wait<0>();
  __syncthreads();

  const int n_strips = (Ls + 15) >> 4;                      // per sample
  __nv_bfloat16* Qs = Qw + warp * 16 * kPitch;
  __nv_bfloat16* dOs = dOw + warp * 16 * kPitch;
  while (true) {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(counter, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= n_strips * nb) break;
    const int si = ticket % nb;
    const int b = b0 + si;
    const int soff = si * Ls;                               // shared-memory row offset of this sample's own keys
    const int q0 = Lc + (n_strips - 1 - ticket / nb) * 16;  // positions
    const int64_t bh = (int64_t)b * H + h;
    stage_strip<HD>(Qs, gbase + (int64_t)b * Ls * ld, ld, q0, L, lane);
    stage_strip<HD>(dOs, dout + (int64_t)b * Ls * D + (int64_t)h * HD, D, q0 - Lc, Ls, lane);
    __syncwarp();
    const int row_a = q0 + g, row_b = row_a + 8;
    const float lse_a = row_a < L ? lse[bh * Ls + row_a - Lc] * 1.4426950408889634f : INFINITY;
    const float lse_b = row_b < L ? lse[bh * Ls + row_b - Lc] * 1.4426950408889634f : INFINITY;
    const float del_a = row_a < L ? delta[bh * Ls + row_a - Lc] : 0.f;
    const float del_b = row_b < L ? delta[bh * Ls + row_b - Lc] : 0.f;
    float dq[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    const int last_row = min(q0 + 15, L - 1);
    for (int j0 = 0; j0 <= last_row; j0 += 64) {
      const int g_hi = min(4, (last_row - j0) / 16 + 1);
      float s[8][4], dp[8][4];
      mma_a_bt<HD>(s, Qs, 0, Ks + j0 * kPitch, lane, 0, g_hi, Lc - j0, soff);
      mma_a_bt<HD>(dp, dOs, 0, Vs + j0 * kPitch, lane, 0, g_hi, Lc - j0, soff);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = j0 + nt * 8 + tq * 2 + (e & 1);
          const int row = (e < 2) ? row_a : row_b;
          const float pv = (col > row || col >= L) ? 0.f : exp2f(s[nt][e] * scale_log2e - ((e < 2) ? lse_a : lse_b));
          float dpv = dp[nt][e];
          if (drop_thresh != 0)
            dpv = dropout_keep(drop_seed, ((uint64_t)bh * L + row) * (uint64_t)L + col, drop_thresh) ? dpv * drop_scale : 0.f;
          s[nt][e] = pv * (dpv - ((e < 2) ? del_a : del_b)) * scale;
        }
      }
      mma_p_b<HD>(dq, s, Ks + j0 * kPitch, lane, 0, g_hi, Lc - j0, soff);
    }
    // dqkv rows are the samples' own rows: row of position p = b*Ls + p - Lc (store_grad_rows indexes by position)
    store_grad_rows<HD>(dq, dqkv + ((int64_t)b * Ls - Lc) * ld + (int64_t)h * HD, ld, row_a, row_b, L, tq, rope_cos, rope_sin);
    __syncwarp();
  }
}

template <int HD>
__global__ void __launch_bounds__(kSeqThreads)
attn_bwd_dkv_seq_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rope_cos,
                        const float* __restrict__ rope_sin, const __nv_bfloat16* __restrict__ dout,
                        const float* __restrict__ lse, const float* __restrict__ delta,
                        __nv_bfloat16* __restrict__ dqkv, int Bp, int L, int spc, int H, float scale,
                        uint32_t drop_thresh, float drop_scale, uint64_t drop_seed) {
  pdl_wait();
  pdl_trigger();
  // drop_thresh != 0: attention-probability dropout of the forward: dV = (mask * P)^T dO, dS from the masked dP.
  // A CTA serves spc consecutive samples of one head (short own-token runs would otherwise leave most of the 8
  // warps without a key strip): sample i keeps its Q / dO / lse / delta rows at [i*Lp, i*Lp + L).
  constexpr int kPitch = HD + 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const int Lp = (L + 63) & ~63;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* dOs = Qs + (size_t)spc * Lp * kPitch;
  __nv_bfloat16* Kw = dOs + (size_t)spc * Lp * kPitch;     // [8 warps][16][kPitch]
  __nv_bfloat16* Vw = Kw + 8 * 16 * kPitch;
  float* lse_s = reinterpret_cast<float*>(Vw + 8 * 16 * kPitch);
  float* del_s = lse_s + spc * Lp;
  int* counter = reinterpret_cast<int*>(del_s + spc * Lp);

  const int grp = blockIdx.x / H, h = blockIdx.x - grp * H;
  const int b0 = grp * spc, nb = min(spc, Bp - b0);
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const float scale_log2e = scale * 1.4426950408889634f;

  for (int si = 0; si < nb; ++si) {
    const int64_t b = b0 + si;
    stage_rows_async<HD>(Qs + (size_t)si * Lp * kPitch, qkv + b * L * ld + (int64_t)h * HD, ld, L, Lp, threadIdx.x, kSeqThreads);
    stage_rows_async<HD>(dOs + (size_t)si * Lp * kPitch, dout + b * L * D + (int64_t)h * HD, D, L, Lp, threadIdx.x, kSeqThreads);
  }
  cp_async_commit();
  for (int i = threadIdx.x; i < nb * Lp; i += kSeqThreads) {
    const int si = i / Lp, r = i - si * Lp;
    const int64_t bh = (int64_t)(b0 + si) * H + h;
    lse_s[i] = r < L ? lse[bh * L + r] * 1.4426950408889634f : INFINITY;
    del_s[i] = r < L ? delta[bh * L + r] : 0.f;
  }
  if (threadIdx.x == 0) *counter = 0;
  cp_async_wait<0>();
  __syncthreads();

  const int n_strips = (L + 15) >> 4;           // per sample
  __nv_bfloat16* Ks = Kw + warp * 16 * kPitch;
  __nv_bfloat16* Vs = Vw + warp * 16 * kPitch;
  while (true) {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(counter, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= n_strips * nb) break;
    const int si = ticket % nb;
    const int k0 = (ticket / nb) * 16;          // earliest keys see the most queries: heaviest first
    const int64_t b = b0 + si;
    const __nv_bfloat16* qbase = qkv + b * L * ld + (int64_t)h * HD;
    const __nv_bfloat16* Qb = Qs + (size_t)si * Lp * kPitch;
    const __nv_bfloat16* dOb = dOs + (size_t)si * Lp * kPitch;
    const float* lse_b = lse_s + si * Lp;
    const float* del_b = del_s + si * Lp;
    stage_strip<HD>(Ks, qbase + D, ld, k0, L, lane);
    stage_strip<HD>(Vs, qbase + 2 * D, ld, k0, L, lane);
    __syncwarp();
    const int key_a = k0 + g, key_b = key_a + 8;
    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int i0 = (k0 / 64) * 64; i0 < L; i0 += 64) {
      // query groups of 16 entirely before this key strip cannot attend to it
      const int g_lo = max(0, (k0 - i0) / 16);
      const int g_hi = min(4, (L - 1 - i0) / 16 + 1);
      float st[8][4], dpt[8][4];                 // rows = keys, cols = queries
      mma_a_bt<HD, false>(st, Ks, 0, Qb + i0 * kPitch, lane, g_lo, g_hi);
      mma_a_bt<HD, false>(dpt, Vs, 0, dOb + i0 * kPitch, lane, g_lo, g_hi);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qrow = i0 + nt * 8 + tq * 2 + (e & 1);
          const int key = (e < 2) ? key_a : key_b;
          const bool dead = key > qrow || key >= L || qrow >= L || (nt >> 1) < g_lo || (nt >> 1) >= g_hi;
          const float pv = dead ? 0.f : exp2f(st[nt][e] * scale_log2e - lse_b[min(qrow, Lp - 1)]);
          float pd = pv, dpv = dpt[nt][e];
          if (drop_thresh != 0 && !dead) {
            const bool keep = dropout_keep(drop_seed, (((uint64_t)b * H + h) * L + qrow) * (uint64_t)L + key, drop_thresh);
            pd = keep ? pv * drop_scale : 0.f;
            dpv = keep ? dpv * drop_scale : 0.f;
          }
          st[nt][e] = pd;
          dpt[nt][e] = dead ? 0.f : pv * (dpv - del_b[min(qrow, Lp - 1)]) * scale;
        }
      }
      mma_p_b<HD>(dv, st, dOb + i0 * kPitch, lane, g_lo, g_hi);
      mma_p_b<HD>(dk, dpt, Qb + i0 * kPitch, lane, g_lo, g_hi);
    }
    __nv_bfloat16* dbase = dqkv + b * L * ld + (int64_t)h * HD;
    store_grad_rows<HD>(dk, dbase + D, ld, key_a, key_b, L, tq, rope_cos, rope_sin);
    store_grad_rows<HD>(dv, dbase + 2 * D, ld, key_a, key_b, L, tq, nullptr, nullptr);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-prefix layout, frozen backbone: the whole backward of the samples' own tokens in ONE kernel.  A CTA
// serves spc samples of one head and keeps everything resident — K and V of the prefix and of its samples, Q and
// dO of its samples — staged once with cp.async.  It first computes delta = rowsum(dO * O) for its rows (the
// separate delta pass disappears), then its 8 warps pull jobs from a shared counter: the dQ strips (16 own queries
// against prefix + own keys), then the dK/dV strips (16 own keys against the sample's own queries).  Operands
// come straight from the resident arrays: no per-warp staging, no second launch, no second staging of K/V/Q/dO.
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kSeqThreads)
attn_bwd_own_fused_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rope_cos,
                          const float* __restrict__ rope_sin, const __nv_bfloat16* __restrict__ out_own,
                          const __nv_bfloat16* __restrict__ dout_own, const float* __restrict__ lse_own,
                          __nv_bfloat16* __restrict__ dqkv_own, int Bp, int L, int Lc, int spc, int rows_alloc, int H,
                          float scale) {
  pdl_wait();
  pdl_trigger();
  constexpr int kPitch = HD + 8;
  constexpr int kVec = HD / 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const int Ls = L - Lc;
  const int Lq = (Ls + 63) & ~63;                           // rows kept per sample for Q / dO (zero padded)
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Vs = Ks + (size_t)rows_alloc * kPitch;
  __nv_bfloat16* Qs = Vs + (size_t)rows_alloc * kPitch;     // [spc][Lq][kPitch]
  __nv_bfloat16* dOs = Qs + (size_t)spc * Lq * kPitch;
  float* lse_s = reinterpret_cast<float*>(dOs + (size_t)spc * Lq * kPitch);   // [spc][Lq], pre-multiplied by log2(e)
  float* del_s = lse_s + spc * Lq;
  int* counter = reinterpret_cast<int*>(del_s + spc * Lq);

  const int grp = blockIdx.x / H, h = blockIdx.x - grp * H;
  const int b0 = grp * spc, nb = min(spc, Bp - b0);
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const __nv_bfloat16* gbase = qkv + (int64_t)h * HD;      // row of (sample b, position p): p (p < Lc) or p + b*Ls
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const float scale_log2e = scale * 1.4426950408889634f;

  // ---- stage K, V (prefix + own rows of the nb samples: one contiguous range of global rows after the prefix)
  const int rows_used = Lc + nb * Ls;
  for (int i = threadIdx.x; i < rows_alloc * kVec; i += kSeqThreads) {
    const int r = i / kVec, c = (i - r * kVec) * 8;
    if (r < rows_used) {
      const __nv_bfloat16* src = gbase + ((int64_t)r + (r < Lc ? 0 : (int64_t)b0 * Ls)) * ld + c;
      cp_async16(smem_u32(Ks + r * kPitch + c), src + D);
      cp_async16(smem_u32(Vs + r * kPitch + c), src + 2 * D);
    } else {
      *reinterpret_cast<uint4*>(Ks + r * kPitch + c) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Vs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
    }
  }
  // ---- stage Q, dO of the samples' own tokens
  for (int i = threadIdx.x; i < spc * Lq * kVec; i += kSeqThreads) {
    const int rr = i / kVec, c = (i - rr * kVec) * 8;
    const int si = rr / Lq, t = rr - si * Lq;
    if (si < nb && t < Ls) {
      const int64_t own_row = (int64_t)(b0 + si) * Ls + t;                  // row among the own rows
      cp_async16(smem_u32(Qs + rr * kPitch + c), gbase + (own_row + Lc) * ld + c);
      cp_async16(smem_u32(dOs + rr * kPitch + c), dout_own + own_row * D + (int64_t)h * HD + c);
    } else {
      *reinterpret_cast<uint4*>(Qs + rr * kPitch + c) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dOs + rr * kPitch + c) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_commit();
  if (threadIdx.x == 0) *counter = 0;
  cp_async_wait<0>();
  __syncthreads();
  // ---- delta = rowsum(dO * O) and lse of the own rows: two threads per row, each with its half row of O in flight
  // as independent 16-byte loads (one global round trip for the whole CTA)
  for (int idx = threadIdx.x; idx < 2 * spc * Lq; idx += kSeqThreads) {
    const int rr = idx >> 1, part = idx & 1;
    const int si = rr / Lq, t = rr - si * Lq;
    const bool live = si < nb && t < Ls;
    float acc = 0.f;
    if (live) {
      const int64_t own_row = (int64_t)(b0 + si) * Ls + t;
      const uint4* op = reinterpret_cast<const uint4*>(out_own + own_row * D + (int64_t)h * HD + part * (HD / 2));
      const uint4* dp = reinterpret_cast<const uint4*>(dOs + rr * kPitch + part * (HD / 2));
      uint4 ov[HD / 16];
#pragma unroll
      for (int i = 0; i < HD / 16; ++i) ov[i] = __ldg(op + i);
#pragma unroll
      for (int i = 0; i < HD / 16; ++i) {
        const uint4 d = dp[i];
        const uint32_t aw[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += bf16_lo(aw[q]) * bf16_lo(dw[q]) + bf16_hi(aw[q]) * bf16_hi(dw[q]);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (part == 0) {
      del_s[rr] = acc;
      lse_s[rr] = live ? lse_own[((int64_t)(b0 + si) * H + h) * Ls + t] * 1.4426950408889634f : INFINITY;
    }
  }
  __syncthreads();

  const int n_strips = (Ls + 15) >> 4;                      // per sample, for both job kinds
  const int n_dq = n_strips * nb;
  while (true) {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(counter, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= 2 * n_dq) break;
    if (ticket < n_dq) {
      // ------------------------------------------------------------ dQ strip: 16 own queries x (prefix + own keys)
      const int si = ticket % nb;
      const int t0 = (n_strips - 1 - ticket / nb) * 16;     // heaviest (latest) strips first
      const int q0 = Lc + t0;                               // positions
      const int soff = si * Ls;
      const __nv_bfloat16* Qb = Qs + (size_t)si * Lq * kPitch;
      const __nv_bfloat16* dOb = dOs + (size_t)si * Lq * kPitch;
      const int row_a = q0 + g, row_b = row_a + 8;
      const float lse_a = row_a < L ? lse_s[si * Lq + t0 + g] : INFINITY;
      const float lse_b = row_b < L ? lse_s[si * Lq + t0 + g + 8] : INFINITY;
      const float del_a = row_a < L ? del_s[si * Lq + t0 + g] : 0.f;
      const float del_b = row_b < L ? del_s[si * Lq + t0 + g + 8] : 0.f;
      float dq[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
      const int last_row = min(q0 + 15, L - 1);
      for (int j0 = 0; j0 <= last_row; j0 += 64) {
        const int g_hi = min(4, (last_row - j0) / 16 + 1);
        float s[8][4], dp[8][4];
        mma_a_bt<HD>(s, Qb, t0, Ks + j0 * kPitch, lane, 0, g_hi, Lc - j0, soff);
        mma_a_bt<HD>(dp, dOb, t0, Vs + j0 * kPitch, lane, 0, g_hi, Lc - j0, soff);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = j0 + nt * 8 + tq * 2 + (e & 1);
            const int row = (e < 2) ? row_a : row_b;
            const float pv = (col > row || col >= L) ? 0.f : exp2f(s[nt][e] * scale_log2e - ((e < 2) ? lse_a : lse_b));
            s[nt][e] = pv * (dp[nt][e] - ((e < 2) ? del_a : del_b)) * scale;
          }
        }
        mma_p_b<HD>(dq, s, Ks + j0 * kPitch, lane, 0, g_hi, Lc - j0, soff);
      }
      store_grad_rows<HD>(dq, dqkv_own + ((int64_t)(b0 + si) * Ls - Lc) * ld + (int64_t)h * HD, ld, row_a, row_b, L, tq,
                          rope_cos, rope_sin);
    } else {
      // ------------------------------------------------------------ dK / dV strip: 16 own keys x the sample's own queries
      const int tk = ticket - n_dq;
      const int si = tk % nb;
      const int k0 = (tk / nb) * 16;                        // own-token index of the first key; earliest = heaviest
      const __nv_bfloat16* Qb = Qs + (size_t)si * Lq * kPitch;
      const __nv_bfloat16* dOb = dOs + (size_t)si * Lq * kPitch;
      const float* lse_b = lse_s + si * Lq;
      const float* del_b = del_s + si * Lq;
      const int krow = Lc + si * Ls + k0;                   // shared-memory row of the strip's first key
      const int key_a = k0 + g, key_b = key_a + 8;
      float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
      }
      for (int i0 = (k0 / 64) * 64; i0 < Ls; i0 += 64) {
        const int g_lo = max(0, (k0 - i0) / 16);
        const int g_hi = min(4, (Ls - 1 - i0) / 16 + 1);
        float st[8][4], dpt[8][4];                          // rows = keys, cols = queries
        mma_a_bt<HD, false>(st, Ks, krow, Qb + i0 * kPitch, lane, g_lo, g_hi);
        mma_a_bt<HD, false>(dpt, Vs, krow, dOb + i0 * kPitch, lane, g_lo, g_hi);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int qrow = i0 + nt * 8 + tq * 2 + (e & 1);
            const int key = (e < 2) ? key_a : key_b;
            const bool dead = key > qrow || key >= Ls || qrow >= Ls || (nt >> 1) < g_lo || (nt >> 1) >= g_hi;
            const float pv = dead ? 0.f : exp2f(st[nt][e] * scale_log2e - lse_b[min(qrow, Lq - 1)]);
            st[nt][e] = pv;
            dpt[nt][e] = dead ? 0.f : pv * (dpt[nt][e] - del_b[min(qrow, Lq - 1)]) * scale;
          }
        }
        mma_p_b<HD>(dv, st, dOb + i0 * kPitch, lane, g_lo, g_hi);
        mma_p_b<HD>(dk, dpt, Qb + i0 * kPitch, lane, g_lo, g_hi);
      }
      // own rows of dqkv; RoPE position of own token t is Lc + t
      __nv_bfloat16* dbase = dqkv_own + (int64_t)(b0 + si) * Ls * ld + (int64_t)h * HD;
      const int half = HD / 2;
      store_grad_rows<HD>(dk, dbase + D, ld, key_a, key_b, Ls, tq, rope_cos ? rope_cos + (int64_t)Lc * half : nullptr,
                          rope_sin ? rope_sin + (int64_t)Lc * half : nullptr);
      store_grad_rows<HD>(dv, dbase + 2 * D, ld, key_a, key_b, Ls, tq, nullptr, nullptr);
    }
  }
}

template <int HD>
static size_t fused_bwd_smem_bytes(int L, int Lc, int spc) {
  const int Lq = ((L - Lc) + 63) & ~63;
  return (size_t)(2 * seq_rows_alloc(L, Lc, spc) + 2 * spc * Lq) * (HD + 8) * 2 + (size_t)2 * spc * Lq * 4 + 16;
}

// ---------------------------------------------------------------------------------------------
// Shared-prefix layout, dK / dV of the PREFIX keys (needed when something trainable sits inside the backbone:
// LoRA).  Every query of the batch attends to the prefix — the prefix's own Lc queries causally, the Bp*Ls own
// tokens of all samples without a mask — and in the shared-prefix row layout all those queries are simply the
// rows [0, Lc + Bp*Ls) of qkv.  One CTA per (head, block of 128 prefix keys): each of the 8 warps owns a 16-key
// strip (K, V strip staged once) and sweeps over all query rows in chunks of 64 (Q, dO, lse, delta of the next
// chunk are fetched with cp.async while the current one is in the tensor cores).  No atomics.
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kSeqThreads)
attn_bwd_dkv_prefix_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rope_cos,
                           const float* __restrict__ rope_sin, const __nv_bfloat16* __restrict__ dout,
                           const float* __restrict__ lse, const float* __restrict__ delta,
                           __nv_bfloat16* __restrict__ dqkv, int Bp, int Lc, int Ls, int H, float scale) {
  pdl_wait();
  pdl_trigger();
  // lse / delta: [H, Lc] for the prefix rows followed by [Bp, H, Ls] for the own rows
  constexpr int kPitch = HD + 8;
  constexpr int kVec = HD / 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(attn_smem);     // [2 stages][64][kPitch]
  __nv_bfloat16* dOs = Qs + 2 * 64 * kPitch;                            // [2 stages][64][kPitch]
  __nv_bfloat16* Kw = dOs + 2 * 64 * kPitch;                            // [8 warps][16][kPitch]
  __nv_bfloat16* Vw = Kw + 8 * 16 * kPitch;
  float* lse_s = reinterpret_cast<float*>(Vw + 8 * 16 * kPitch);        // [2][64]
  float* del_s = lse_s + 2 * 64;

  const int kblocks = (Lc + 127) / 128;
  const int h = blockIdx.x / kblocks, kb = blockIdx.x - h * kblocks;
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const int M = Lc + Bp * Ls;
  const __nv_bfloat16* qbase = qkv + (int64_t)h * HD;
  const __nv_bfloat16* dobase = dout + (int64_t)h * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const float scale_log2e = scale * 1.4426950408889634f;
  const float* lse_own = lse + (int64_t)H * Lc;
  const float* del_own = delta + (int64_t)H * Lc;

  auto stage_chunk = [&](int stage, int i0) {
    __nv_bfloat16* q = Qs + stage * 64 * kPitch;
    __nv_bfloat16* d = dOs + stage * 64 * kPitch;
    for (int i = threadIdx.x; i < 64 * kVec; i += kSeqThreads) {
      const int r = i / kVec, c = (i - r * kVec) * 8;
      if (i0 + r < M) {
        cp_async16(smem_u32(q + r * kPitch + c), qbase + (int64_t)(i0 + r) * ld + c);
        cp_async16(smem_u32(d + r * kPitch + c), dobase + (int64_t)(i0 + r) * D + c);
      } else {
        *reinterpret_cast<uint4*>(q + r * kPitch + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(d + r * kPitch + c) = make_uint4(0, 0, 0, 0);
      }
    }
    if (threadIdx.x < 64) {
      const int row = i0 + threadIdx.x;
      float l = INFINITY, dl = 0.f;
      if (row < Lc) {
        l = lse[(int64_t)h * Lc + row]; dl = delta[(int64_t)h * Lc + row];
      } else if (row < M) {
        const int o = row - Lc, b = o / Ls, t = o - b * Ls;
        l = lse_own[((int64_t)b * H + h) * Ls + t]; dl = del_own[((int64_t)b * H + h) * Ls + t];
      }
      lse_s[stage * 64 + threadIdx.x] = l * 1.4426950408889634f;
      del_s[stage * 64 + threadIdx.x] = dl;
    }
  };

  const int k0 = kb * 128 + warp * 16;            // this warp's key strip (positions = rows)
  __nv_bfloat16* Ks = Kw + warp * 16 * kPitch;
  __nv_bfloat16* Vs = Vw + warp * 16 * kPitch;
  stage_strip<HD>(Ks, qbase + D, ld, k0, Lc, lane);
  stage_strip<HD>(Vs, qbase + 2 * D, ld, k0, Lc, lane);
  // prefix queries before this CTA's first key cannot attend to it: start at the 64-row chunk holding that key
  const int i_begin = (kb * 128 / 64) * 64;
  stage_chunk(0, i_begin);
  cp_async_commit();

  const int key_a = k0 + g, key_b = key_a + 8;
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  int stage = 0;
  for (int i0 = i_begin; i0 < M; i0 += 64, stage ^= 1) {
    if (i0 + 64 < M) stage_chunk(stage ^ 1, i0 + 64);     // (that buffer was released by the barrier ending the previous turn)
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (k0 < Lc) {
      const __nv_bfloat16* Qb = Qs + stage * 64 * kPitch;
      const __nv_bfloat16* dOb = dOs + stage * 64 * kPitch;
      const float* lse_b = lse_s + stage * 64;
      const float* del_b = del_s + stage * 64;
      const int g_hi = min(4, (M - 1 - i0) / 16 + 1);
      float st[8][4], dpt[8][4];                 // rows = keys, cols = queries
      mma_a_bt<HD, false>(st, Ks, 0, Qb, lane, 0, g_hi);
      mma_a_bt<HD, false>(dpt, Vs, 0, dOb, lane, 0, g_hi);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = nt * 8 + tq * 2 + (e & 1);
          const int qrow = i0 + qi;
          const int key = (e < 2) ? key_a : key_b;
          // prefix queries are causal, own tokens (rows >= Lc, positions >= Lc) see every prefix key
          const bool dead = key >= Lc || qrow >= M || (qrow < Lc && key > qrow) || (nt >> 1) >= g_hi;
          const float pv = dead ? 0.f : exp2f(st[nt][e] * scale_log2e - lse_b[qi]);
          st[nt][e] = pv;
          dpt[nt][e] = dead ? 0.f : pv * (dpt[nt][e] - del_b[qi]) * scale;
        }
      }
      mma_p_b<HD>(dv, st, dOb, lane, 0, g_hi);
      mma_p_b<HD>(dk, dpt, Qb, lane, 0, g_hi);
    }
    __syncthreads();                              // everyone is done with this stage before it is refilled
  }
  if (k0 < Lc) {
    __nv_bfloat16* dbase = dqkv + (int64_t)h * HD;
    store_grad_rows<HD>(dk, dbase + D, ld, key_a, key_b, Lc, tq, rope_cos, rope_sin);
    store_grad_rows<HD>(dv, dbase + 2 * D, ld, key_a, key_b, Lc, tq, nullptr, nullptr);
  }
}

template <int HD>
static size_t dkv_prefix_smem_bytes() {
  return (size_t)(4 * 64 + 2 * 8 * 16) * (HD + 8) * 2 + (size_t)4 * 64 * 4;
}

static int g_fused_bwd = -1;
static bool fused_bwd_enabled() {        // MTS_ATTN_FUSED_BWD=0: separate delta / dQ / dK,dV kernels
  if (g_fused_bwd < 0) {
    const char* e = getenv("MTS_ATTN_FUSED_BWD");
    g_fused_bwd = (e && e[0] == '0') ? 0 : 1;
  }
  return g_fused_bwd == 1;
}

template <int HD>
static size_t seq_bwd_smem_bytes(int L, int spc = 1) {
  const int Lp = spc * ((L + 63) & ~63);
  return (size_t)(2 * Lp + 2 * 8 * 16) * (HD + 8) * 2 + (size_t)2 * Lp * 4 + 16;
}
// samples per CTA of the dK/dV kernel: enough 16-key strips for the 8 warps, as far as shared memory allows
template <int HD>
static int seq_dkv_spc(int Bp, int L) {
  const int n_strips = (L + 15) / 16;
  const int want = std::min(Bp, (8 + n_strips - 1) / n_strips);
  int spc = 1;
  while (spc < want && seq_bwd_smem_bytes<HD>(L, spc + 1) <= 220 * 1024) ++spc;
  return spc;
}
// dQ kernel: K/V rows of spc samples sharing Lc of their L positions + Q / dO strip staging
template <int HD>
static size_t seq_dq_smem_bytes(int L, int Lc, int spc) {
  return (size_t)(2 * seq_rows_alloc(L, Lc, spc) + 2 * 8 * 16) * (HD + 8) * 2 + 16;
}

template <int HD>
static int launch_attn_bwd(const uint16_t* qkv, const float* rc, const float* rs, const uint16_t* out,
                           const uint16_t* dout, const float* lse, float* delta, uint16_t* dqkv, int Bp,
                           int L, int H, float scale, int pre_roped, cudaStream_t stream, uint32_t drop_thresh = 0,
                           float drop_scale = 1.0f, uint64_t drop_seed = 0) {
  constexpr int kSmemQ = 4 * 64 * (HD + 8) * 2;
  constexpr int kSmemKV = kSmemQ + 2 * 64 * 4;
  auto kq = attn_bwd_dq_kernel<HD>;
  auto kkv = attn_bwd_dkv_kernel<HD>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemQ);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemKV);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn bwd)", e);
    attr_done = true;
  }
  const int64_t nwarps = (int64_t)Bp * L * H;
  LAUNCH_PDL(attn_bwd_delta_kernel, (int)((nwarps + 7) / 8), 256, 0, stream,
      
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta, L, H, HD, nwarps);
  count_launch();
  int rc_ = check_launch("attn_bwd_delta_kernel");
  if (rc_) return rc_;
  if (pre_roped || rc == nullptr) {
    // q / k in `qkv` are final (rotated or rope-free): short sequences take the sequence-resident kernels
    if (seq_bwd_smem_bytes<HD>(L) <= 220 * 1024) {
      auto sq = attn_bwd_dq_seq_kernel<HD>;
      auto skv = attn_bwd_dkv_seq_kernel<HD>;
      static bool seq_attr = false;
      if (!seq_attr) {
        cudaError_t e = cudaFuncSetAttribute(sq, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(skv, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn bwd seq)", e);
        seq_attr = true;
      }
      const size_t smem = seq_bwd_smem_bytes<HD>(L);
      LAUNCH_PDL(sq, Bp * H, kSeqThreads, smem, stream,
      
          reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(dout), lse,
          delta, reinterpret_cast<__nv_bfloat16*>(dqkv), Bp, L, 0, 1, seq_rows_alloc(L, 0, 1), H, scale, drop_thresh,
          drop_scale, drop_seed);
      count_launch();
      rc_ = check_launch("attn_bwd_dq_seq_kernel");
      if (rc_) return rc_;
      const int spc = seq_dkv_spc<HD>(Bp, L);
      LAUNCH_PDL(skv, ((Bp + spc - 1) / spc) * H, kSeqThreads, seq_bwd_smem_bytes<HD>(L, spc), stream,
      
          reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(dout), lse,
          delta, reinterpret_cast<__nv_bfloat16*>(dqkv), Bp, L, spc, H, scale, drop_thresh, drop_scale, drop_seed);
      count_launch();
      return check_launch("attn_bwd_dkv_seq_kernel");
    }
  }
  if (drop_thresh != 0)
    return set_error(MTS_ERR_UNSUPPORTED, "attention-probability dropout needs pre-rotated q / k and a sequence that fits "
                                          "the sequence-resident kernels (%d positions do not)", L);
  const int64_t grid_l = (int64_t)((L + 63) / 64) * Bp * H;
  if (grid_l > 0x7fffffffLL) return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_bwd: grid too large");
  kq<<<(int)grid_l, kAttnThreads, kSmemQ, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta,
      reinterpret_cast<__nv_bfloat16*>(dqkv), L, H, scale, pre_roped);
  count_launch();
  rc_ = check_launch("attn_bwd_dq_kernel");
  if (rc_) return rc_;
  kkv<<<(int)grid_l, kAttnThreads, kSmemKV, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta,
      reinterpret_cast<__nv_bfloat16*>(dqkv), L, H, scale, pre_roped);
  count_launch();
  return check_launch("attn_bwd_dkv_kernel");
}

// ---------------------------------------------------------------------------------------------
// Shared-prefix variants: the first Lc positions of every sample are the same prompt tokens (static dataset /
// task prompt, models/medtsllm.py:386-439) and, the mask being causal, their states at every layer are the same
// for all samples.  They are kept ONCE in rows [0, Lc); sample b owns rows Lc + b*Ls + [0, Ls).
// ---------------------------------------------------------------------------------------------
template <int HD>
static int launch_attn_shared(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int Lc, int Ls, int H,
                              float scale, cudaStream_t stream) {
  const int L = Lc + Ls;
  if (attn_tc_eligible(L, Lc, HD, Bp, H)) return launch_attn_tc(qkv, out, lse, Bp, L, Lc, H, HD, scale, stream);
  if (seq_smem_bytes<HD>(L) > 220 * 1024)
    return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_shared: %d positions do not fit in shared memory", L);
  // one launch: H CTAs for the prefix as an ordinary sequence of Lc positions, the rest for the samples' own tokens
  return launch_attn_seq<HD>(qkv, out, lse, Bp, L, Lc, H, scale, stream);
}

template <int HD>
static int launch_attn_shared_bwd(const uint16_t* qkv, const float* rc, const float* rs, const uint16_t* out_own,
                                  const uint16_t* dout_own, const float* lse_own, float* delta, uint16_t* dqkv_own,
                                  int Bp, int Lc, int Ls, int H, float scale, cudaStream_t stream,
                                  bool delta_needed_later = false) {
  // delta_needed_later: the caller reads `delta` of the own rows afterwards (full backward: dK/dV of the prefix keys),
  // so the fused kernel — which keeps delta in shared memory only — cannot be used
  const int L = Lc + Ls;
  const int D = H * HD;
  auto sq = attn_bwd_dq_seq_kernel<HD>;
  auto skv = attn_bwd_dkv_seq_kernel<HD>;
  static bool seq_attr = false;      // (first: the full backward launches these kernels for the prefix rows whatever route the own rows take)
  if (!seq_attr) {
    cudaError_t e = cudaFuncSetAttribute(sq, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(skv, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn bwd seq)", e);
    seq_attr = true;
  }
  // tensor-memory kernel (attention_tc.cu); it leaves delta of the own rows behind when the full backward needs it
  if (attn_tc_bwd_eligible(Lc, Ls, HD, Bp, H, rc, rs))
    return launch_attn_bwd_tc(qkv, rc, rs, out_own, dout_own, lse_own, delta_needed_later ? delta : nullptr, dqkv_own, Bp, Lc,
                              Ls, H, HD, scale, stream);
  if (seq_bwd_smem_bytes<HD>(L) > 220 * 1024)
    return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_shared_bwd: %d positions do not fit in shared memory", L);
  // everything of one CTA's samples resident at once?  then one fused kernel does delta, dQ and dK/dV
  if (!delta_needed_later && fused_bwd_enabled() && fused_bwd_smem_bytes<HD>(L, Lc, 1) <= 220 * 1024) {
    int spc = 1;
    const int n_strips = (Ls + 15) / 16;
    const int want = std::min(Bp, (8 + n_strips - 1) / n_strips);      // 8 dQ strips (+ 8 lighter dK/dV strips) for 8 warps
    while (spc < want && fused_bwd_smem_bytes<HD>(L, Lc, spc + 1) <= 220 * 1024) ++spc;
    auto kf = attn_bwd_own_fused_kernel<HD>;
    static bool fattr = false;
    if (!fattr) {
      cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn bwd fused)", e);
      fattr = true;
    }
    LAUNCH_PDL(kf, ((Bp + spc - 1) / spc) * H, kSeqThreads, fused_bwd_smem_bytes<HD>(L, Lc, spc), stream,
               reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(out_own),
               reinterpret_cast<const __nv_bfloat16*>(dout_own), lse_own, reinterpret_cast<__nv_bfloat16*>(dqkv_own), Bp, L,
               Lc, spc, seq_rows_alloc(L, Lc, spc), H, scale);
    count_launch();
    return check_launch("attn_bwd_own_fused_kernel");
  }
  const int64_t nwarps = (int64_t)Bp * Ls * H;
  LAUNCH_PDL(attn_bwd_delta_kernel, (int)((nwarps + 7) / 8), 256, 0, stream,
      
      reinterpret_cast<const __nv_bfloat16*>(out_own), reinterpret_cast<const __nv_bfloat16*>(dout_own), delta, Ls, H, HD,
      nwarps);
  count_launch();
  int rc_ = check_launch("attn_bwd_delta_kernel");
  if (rc_) return rc_;
  // dQ of the own tokens: keys = prefix + own; spc samples per CTA (enough strips for the 8 warps, if they fit)
  int spc = 1;
  {
    const int n_strips = (Ls + 15) / 16;
    const int want = std::min(Bp, (8 + n_strips - 1) / n_strips);
    while (spc < want && seq_dq_smem_bytes<HD>(L, Lc, spc + 1) <= 220 * 1024) ++spc;
  }
  LAUNCH_PDL(sq, ((Bp + spc - 1) / spc) * H, kSeqThreads, seq_dq_smem_bytes<HD>(L, Lc, spc), stream,
      
      reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(dout_own), lse_own,
      delta, reinterpret_cast<__nv_bfloat16*>(dqkv_own), Bp, L, Lc, spc, seq_rows_alloc(L, Lc, spc), H, scale, 0u, 1.0f, (uint64_t)0);
  count_launch();
  rc_ = check_launch("attn_bwd_dq_seq_kernel");
  if (rc_) return rc_;
  // dK / dV of the own tokens only (queries = own tokens; the log-sum-exp already covers the prefix keys):
  // the plain kernel on the own rows, RoPE tables shifted to position Lc
  const int half = HD / 2;
  const int spc_kv = seq_dkv_spc<HD>(Bp, Ls);
  LAUNCH_PDL(skv, ((Bp + spc_kv - 1) / spc_kv) * H, kSeqThreads, seq_bwd_smem_bytes<HD>(Ls, spc_kv), stream,
      
      reinterpret_cast<const __nv_bfloat16*>(qkv) + (int64_t)Lc * 3 * D, rc ? rc + (int64_t)Lc * half : nullptr,
      rs ? rs + (int64_t)Lc * half : nullptr, reinterpret_cast<const __nv_bfloat16*>(dout_own), lse_own, delta,
      reinterpret_cast<__nv_bfloat16*>(dqkv_own), Bp, Ls, spc_kv, H, scale, 0u, 1.0f, (uint64_t)0);
  count_launch();
  return check_launch("attn_bwd_dkv_seq_kernel");
}

// Full backward on the shared-prefix layout: gradients for the prefix rows too (out / dout / dqkv hold ALL rows,
// lse / delta = [H, Lc] followed by [Bp, H, Ls]).
template <int HD>
static int launch_attn_shared_bwd_full(const uint16_t* qkv, const float* rc, const float* rs, const uint16_t* out,
                                       const uint16_t* dout, const float* lse, float* delta, uint16_t* dqkv, int Bp,
                                       int Lc, int Ls, int H, float scale, cudaStream_t stream) {
  const int D = H * HD;
  // own rows: dQ (prefix + own keys) and dK / dV of the own keys
  int rc_ = launch_attn_shared_bwd<HD>(qkv, rc, rs, out + (int64_t)Lc * D, dout + (int64_t)Lc * D, lse + (int64_t)H * Lc,
                                       delta + (int64_t)H * Lc, dqkv + (int64_t)Lc * 3 * D, Bp, Lc, Ls, H, scale, stream,
                                       /*delta_needed_later=*/true);
  if (rc_) return rc_;
  // prefix rows: delta, then dQ of the prefix as one causal sequence of Lc positions
  const int64_t nwarps = (int64_t)Lc * H;
  LAUNCH_PDL(attn_bwd_delta_kernel, (int)((nwarps + 7) / 8), 256, 0, stream,
      
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta, Lc, H, HD, nwarps);
  count_launch();
  rc_ = check_launch("attn_bwd_delta_kernel");
  if (rc_) return rc_;
  auto sq = attn_bwd_dq_seq_kernel<HD>;
  LAUNCH_PDL(sq, H, kSeqThreads, seq_dq_smem_bytes<HD>(Lc, 0, 1), stream,
      
      reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta,
      reinterpret_cast<__nv_bfloat16*>(dqkv), 1, Lc, 0, 1, seq_rows_alloc(Lc, 0, 1), H, scale, 0u, 1.0f, (uint64_t)0);
  count_launch();
  rc_ = check_launch("attn_bwd_dq_seq_kernel");
  if (rc_) return rc_;
  // dK / dV of the prefix keys: every query row of the batch contributes
  auto kp = attn_bwd_dkv_prefix_kernel<HD>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkv_prefix_smem_bytes<HD>());
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn bwd prefix)", e);
    attr = true;
  }
  LAUNCH_PDL(kp, H * ((Lc + 127) / 128), kSeqThreads, dkv_prefix_smem_bytes<HD>(), stream,
      
      reinterpret_cast<const __nv_bfloat16*>(qkv), rc, rs, reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta,
      reinterpret_cast<__nv_bfloat16*>(dqkv), Bp, Lc, Ls, H, scale);
  count_launch();
  return check_launch("attn_bwd_dkv_prefix_kernel");
}

}  // namespace mts

using namespace mts;

extern "C" int mts_attn_causal_shared_bwd_full(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                                               const uint16_t* out, const uint16_t* dout, const float* lse, float* delta,
                                               uint16_t* dqkv, int Bp, int Lc, int Ls, int H, int hd, float scale,
                                               mts_stream_t s) {
  if (!qkv || !out || !dout || !lse || !delta || !dqkv || Bp <= 0 || Lc <= 0 || Ls <= 0 || H <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared_bwd_full: bad args");
  if ((rope_cos == nullptr) != (rope_sin == nullptr))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared_bwd_full: rope_cos and rope_sin go together");
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(dout) & 15) ||
      (reinterpret_cast<uintptr_t>(dqkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 3))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared_bwd_full: misaligned pointer");
  switch (hd) {
    case 64: return launch_attn_shared_bwd_full<64>(qkv, rope_cos, rope_sin, out, dout, lse, delta, dqkv, Bp, Lc, Ls, H, scale, (cudaStream_t)s);
    case 128: return launch_attn_shared_bwd_full<128>(qkv, rope_cos, rope_sin, out, dout, lse, delta, dqkv, Bp, Lc, Ls, H, scale, (cudaStream_t)s);
    default: return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_shared_bwd_full: head dim %d (supported: 64, 128)", hd);
  }
}

extern "C" int mts_attn_causal_shared(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int Lc, int Ls, int H,
                                      int hd, float scale, mts_stream_t s) {
  if (!qkv || !out || Bp <= 0 || Lc <= 0 || Ls <= 0 || H <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared: bad args");
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared: misaligned pointer");
  switch (hd) {
    case 64: return launch_attn_shared<64>(qkv, out, lse, Bp, Lc, Ls, H, scale, (cudaStream_t)s);
    case 128: return launch_attn_shared<128>(qkv, out, lse, Bp, Lc, Ls, H, scale, (cudaStream_t)s);
    default: return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_shared: head dim %d (supported: 64, 128)", hd);
  }
}

extern "C" int mts_attn_causal_shared_bwd(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                                          const uint16_t* out_own, const uint16_t* dout_own, const float* lse_own,
                                          float* delta, uint16_t* dqkv_own, int Bp, int Lc, int Ls, int H, int hd,
                                          float scale, mts_stream_t s) {
  if (!qkv || !out_own || !dout_own || !lse_own || !delta || !dqkv_own || Bp <= 0 || Lc <= 0 || Ls <= 0 || H <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared_bwd: bad args");
  if ((rope_cos == nullptr) != (rope_sin == nullptr))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared_bwd: rope_cos and rope_sin go together");
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(dout_own) & 15) ||
      (reinterpret_cast<uintptr_t>(dqkv_own) & 15) || (reinterpret_cast<uintptr_t>(out_own) & 3))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_shared_bwd: misaligned pointer");
  switch (hd) {
    case 64: return launch_attn_shared_bwd<64>(qkv, rope_cos, rope_sin, out_own, dout_own, lse_own, delta, dqkv_own, Bp, Lc, Ls, H, scale, (cudaStream_t)s);
    case 128: return launch_attn_shared_bwd<128>(qkv, rope_cos, rope_sin, out_own, dout_own, lse_own, delta, dqkv_own, Bp, Lc, Ls, H, scale, (cudaStream_t)s);
    default: return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_shared_bwd: head dim %d (supported: 64, 128)", hd);
  }
}

extern "C" int mts_attn_causal(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                               uint16_t* out, float* lse, int Bp, int L, int H, int hd, float scale,
                               mts_stream_t s) {
  if (!qkv || !out || Bp <= 0 || L <= 0 || H <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal: bad args");
  if ((rope_cos == nullptr) != (rope_sin == nullptr))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal: rope_cos and rope_sin go together");
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (rope_cos && ((reinterpret_cast<uintptr_t>(rope_cos) & 3) || (reinterpret_cast<uintptr_t>(rope_sin) & 3))))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal: misaligned pointer");
  switch (hd) {
    case 64: return launch_attn<64>(qkv, rope_cos, rope_sin, out, lse, Bp, L, H, scale, (cudaStream_t)s);
    case 128: return launch_attn<128>(qkv, rope_cos, rope_sin, out, lse, Bp, L, H, scale, (cudaStream_t)s);
    default: return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal: head dim %d (supported: 64, 128)", hd);
  }
}

// Attention with dropout on the probabilities (train-mode HF GPT-2: attn_pdrop, HF:models/gpt2/modeling_gpt2.py:67-68;
// Llama: attention_dropout).  q / k pre-rotated, plain row layout, sequence-resident kernels only.
extern "C" int mts_attn_causal_dropout(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int L, int H, int hd,
                                       float scale, float p, uint64_t seed, mts_stream_t stream_) {
  if (!qkv || !out || Bp <= 0 || L <= 0 || H <= 0 || !(p >= 0.f && p < 1.f))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_dropout: bad arguments");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  const float dscale = 1.0f / (1.0f - p);
  if (hd == 64) {
    if (seq_smem_bytes<64>(L) > 220 * 1024) return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_dropout: sequence too long");
    return launch_attn_seq<64>(qkv, out, lse, Bp, L, 0, H, scale, stream, thresh, dscale, seed);
  }
  if (hd == 128) {
    if (seq_smem_bytes<128>(L) > 220 * 1024) return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_dropout: sequence too long");
    return launch_attn_seq<128>(qkv, out, lse, Bp, L, 0, H, scale, stream, thresh, dscale, seed);
  }
  return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_dropout: head dim must be 64 or 128");
}

extern "C" int mts_attn_causal_dropout_bwd(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                                           const uint16_t* out, const uint16_t* dout, const float* lse, float* delta,
                                           uint16_t* dqkv, int Bp, int L, int H, int hd, float scale, float p,
                                           uint64_t seed, mts_stream_t stream_) {
  if (!qkv || !out || !dout || !lse || !delta || !dqkv || Bp <= 0 || L <= 0 || H <= 0 || !(p >= 0.f && p < 1.f))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_dropout_bwd: bad arguments");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  const float dscale = 1.0f / (1.0f - p);
  if (hd == 64) return launch_attn_bwd<64>(qkv, rope_cos, rope_sin, out, dout, lse, delta, dqkv, Bp, L, H, scale, 1, stream, thresh, dscale, seed);
  if (hd == 128) return launch_attn_bwd<128>(qkv, rope_cos, rope_sin, out, dout, lse, delta, dqkv, Bp, L, H, scale, 1, stream, thresh, dscale, seed);
  return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_dropout_bwd: head dim must be 64 or 128");
}

extern "C" int mts_attn_causal_bwd(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                                   const uint16_t* out, const uint16_t* dout, const float* lse,
                                   float* delta, uint16_t* dqkv, int Bp, int L, int H, int hd,
                                   float scale, int pre_roped, mts_stream_t s) {
  if (!qkv || !out || !dout || !lse || !delta || !dqkv || Bp <= 0 || L <= 0 || H <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_bwd: bad args");
  if ((rope_cos == nullptr) != (rope_sin == nullptr))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_bwd: rope_cos and rope_sin go together");
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(dout) & 15) ||
      (reinterpret_cast<uintptr_t>(dqkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 3))
    return set_error(MTS_ERR_INVALID_ARG, "mts_attn_causal_bwd: misaligned pointer");
  switch (hd) {
    case 64: return launch_attn_bwd<64>(qkv, rope_cos, rope_sin, out, dout, lse, delta, dqkv, Bp, L, H, scale, pre_roped, (cudaStream_t)s);
    case 128: return launch_attn_bwd<128>(qkv, rope_cos, rope_sin, out, dout, lse, delta, dqkv, Bp, L, H, scale, pre_roped, (cudaStream_t)s);
    default: return set_error(MTS_ERR_UNSUPPORTED, "mts_attn_causal_bwd: head dim %d (supported: 64, 128)", hd);
  }
}

extern "C" int mts_rope_qk(uint16_t* qkv, const float* rope_cos, const float* rope_sin, int Bp, int L, int H,
                           int hd, mts_stream_t s) {
  if (!qkv || !rope_cos || !rope_sin || Bp <= 0 || L <= 0 || H <= 0 || hd <= 0 || (hd % 16))
    return set_error(MTS_ERR_INVALID_ARG, "mts_rope_qk: bad args (hd must be a multiple of 16)");
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return set_error(MTS_ERR_INVALID_ARG, "mts_rope_qk: misaligned qkv");
  const int64_t rows = (int64_t)Bp * L;
  const int64_t total = rows * 2 * H * (hd / 16);
  int64_t g = (total + 255) / 256;
  if (g > (int64_t)num_sms() * 32) g = (int64_t)num_sms() * 32;
  LAUNCH_PDL(rope_qk_kernel, (int)g, 256, 0, (cudaStream_t)s,
      reinterpret_cast<__nv_bfloat16*>(qkv), rope_cos, rope_sin,
                                                     rows, L, 0, H, hd);
  count_launch();
  return check_launch("rope_qk_kernel");
}

extern "C" int mts_rope_qk_shared(uint16_t* qkv, const float* rope_cos, const float* rope_sin, int Bp, int Lc, int Ls,
                                  int H, int hd, mts_stream_t s) {
  if (!qkv || !rope_cos || !rope_sin || Bp <= 0 || Lc < 0 || Ls <= 0 || H <= 0 || hd <= 0 || (hd % 16))
    return set_error(MTS_ERR_INVALID_ARG, "mts_rope_qk_shared: bad args (hd must be a multiple of 16)");
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return set_error(MTS_ERR_INVALID_ARG, "mts_rope_qk_shared: misaligned qkv");
  const int64_t rows = (int64_t)Lc + (int64_t)Bp * Ls;
  const int64_t total = rows * 2 * H * (hd / 16);
  int64_t g = (total + 255) / 256;
  if (g > (int64_t)num_sms() * 32) g = (int64_t)num_sms() * 32;
  LAUNCH_PDL(rope_qk_kernel, (int)g, 256, 0, (cudaStream_t)s,
      reinterpret_cast<__nv_bfloat16*>(qkv), rope_cos, rope_sin,
                                                     rows, Ls, Lc, H, hd);
  count_launch();
  return check_launch("rope_qk_kernel");
}
