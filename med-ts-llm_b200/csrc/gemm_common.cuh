// gemm_common.cuh — pieces shared by the 1-CTA and the CTA-pair (cta_group::2) tcgen05 GEMM kernels:
// tile constants, kernel parameters, tile rasterisation and the fused epilogue of one 128-row slab.
#pragma once
#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 16;
// warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue set 0, warps 6..9 epilogue set 1.  A warp may only read the
// TMEM lanes of its quarter (warp % 4), so the two sets cover the same rows; when the tile's stores take the direct path
// (GemmParams::direct, whole 32-column chunks) they split the COLUMN chunks between them — the epilogue of a launch with
// one tile per CTA is fully exposed, and it is a latency chain (tcgen05.ld -> residual load -> store) that two warps per
// scheduler overlap.  Otherwise set 1 only hands the accumulator back.
constexpr int kGemmThreads = 320;
constexpr int kNumEpiThreads = 256;
constexpr int kEpiPitch = 36;  // floats per staged row (144 B: 16-byte aligned, bank-conflict free)

struct GemmParams {
  void* d;
  const float* bias;
  int64_t ldd, d_batch_stride;
  int m, n, k, batch;
  int a_batched, b_batched;  // 0: operand shared by all batches
  int bias_axis, d_transposed, d_is_f32, bias_vec, vec_ok;
  const float* c;  // RESID_ADD: residual source, same layout as d
  float alpha;
  // ROPE_QK: rotate-half RoPE on output columns < rope_cols (the q | k sections of a fused qkv projection)
  const float* rope_cos;
  const float* rope_sin;
  int rope_L, rope_hd, rope_cols, rope_prefix;
  // SWIGLU: optional bf16 copy of the gate/up pre-activations (packed column order, row stride ld_aux) that the
  // training path keeps for the backward
  __nv_bfloat16* aux;
  int64_t ld_aux;
  // RESID_ADD with dropout on the projected value: D = C + dropout(alpha * A B^T + bias) — the train-mode residual
  // dropouts of HF GPT-2 (resid_pdrop: HF:models/gpt2/modeling_gpt2.py:233, :243); element (row, col) of the [m, n]
  // result has index row*n + col in the counter-based mask (the backward calls mts_dropout on the [m, n] gradient)
  uint32_t drop_thresh;
  float drop_scale;
  uint64_t drop_seed;
  // stream-K (single-CTA kernel, batch 1): the (tile, k-block) units are dealt out evenly over the CTAs; a CTA whose
  // range STARTS inside a tile parks its raw fp32 partial in sk_ws[slot = its CTA index] (always its first segment)
  // and raises sk_flags[CTA] to sk_epoch; the CTA that holds the tile's FIRST k-blocks (its last segment) adds the
  // partials in ascending CTA order (deterministic) and runs the epilogue.  Contributors never wait: no deadlock, and an
  // owner finds the partials already there when it gets to the end of its own range.
  int streamk;
  int sk_grid;     // stream-K grid size: the SM count, or tiles x s for the even split-K form (every tile cut into s equal
                   // k ranges held by s consecutive CTAs: small tile counts with a deep k, mts_set_option("streamk", 3))
  float* sk_ws;
  int* sk_flags;
  int sk_epoch;
  int precise;     // fp32 operands (evaluation parity modes): accurate expf / tanhf in the activation epilogues
  int split3;      // TF32 kernels: three k sweeps (hi*hi, lo*hi, hi*lo) over the split operands
  int round_tf32;  // fp32 D only: round the stored values to TF32 (they feed a kind::tf32 GEMM next)
  long long* dbg;  // MTS_GEMM_DBG=1 (single-CTA kernel): clock64 stamps of block 0, printed by the launcher (debug)
  int ksplit;      // cluster split-K (single-CTA kernel): clusters of ksplit CTAs share one tile, CTA r of a cluster runs the
                   // r-th k range over the whole tile, sends the column parts it does not own into their owners' shared
                   // memory (the operand ring, idle by then) and runs the epilogue of column part r — no workspace, no
                   // global-memory round trip; 0 / 1 = off
  int direct;      // rows of D (and C) are 32-byte aligned: full 32-column chunks are stored straight from the registers with
                   // 32-byte accesses (a lane's 32 columns are 64 / 128 contiguous bytes) instead of being staged through smem
};
#define GEMM_STAMP(slot) do { if (p.dbg && blockIdx.x == 0) p.dbg[slot] = clock64(); } while (0)

__device__ __forceinline__ float gelu_new_f(float x) {
  // HF:activations.py:65-66  0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));  // |err| ~ 2^-11: below the bf16 output rounding
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// evaluation parity modes (fp32 operands): library-accurate transcendentals instead of the approximate units
__device__ __forceinline__ float gelu_new_precise_f(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float silu_precise_f(float x) { return x / (1.0f + expf(-x)); }

// Tile order inside one batch: groups of kGroupM row-blocks, m fastest inside a group.  The ~148
// tiles in flight then form a roughly square patch (12 x 12 blocks) of the output, so a wave streams
// ~24 operand strips from L2/HBM instead of 48 + 3 with a plain m-fastest order.
constexpr int kGroupM = 12;
__device__ __forceinline__ void tile_coords(int t, int m_blocks, int n_blocks, int& m_blk, int& n_blk) {
  const int per_group = kGroupM * n_blocks;
  const int group = t / per_group;
  const int first_m = group * kGroupM;
  const int gsize = min(kGroupM, m_blocks - first_m);
  const int r = t - group * per_group;
  n_blk = r / gsize;
  m_blk = first_m + (r - n_blk * gsize);
}

// 2 when both epilogue warp sets share a tile's column chunks: every chunk takes the direct store path (no staging buffer,
// which only set 0 owns), i.e. p.direct and a column count that is a multiple of 32.  Warp-uniform, launch-uniform.
template <int EPI>
__device__ __forceinline__ int epilogue_parts(const GemmParams& p) {
  if constexpr (EPI == MTS_EPI_ROPE_QK) return (p.direct && !p.d_is_f32 && (p.n % 32) == 0) ? 2 : 1;
  const int n_store = (EPI == MTS_EPI_SWIGLU) ? p.n / 2 : p.n;
  return (p.direct && !p.d_transposed && (n_store % 32) == 0) ? 2 : 1;
}

// Epilogue of one warp's 32-row slab of a 128 x BN accumulator tile:
//   TMEM -> registers (thread = accumulator row, 32 columns per tcgen05.ld) -> bias / activation ->
//   per-warp padded smem tile -> coalesced 16-byte global accesses.
//   row0  = global row of the slab's first row,   taddr = TMEM address of (slab lane 0, tile column 0)
//   col_off / n_cols: the tile may be a column slice [col_off, col_off + n_cols) of the BN-wide block
//   (tail splitting in the CTA-pair kernel); full tiles pass (0, BN).
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, int b, int row0, int n_blk, uint32_t taddr,
                                              float* stage_buf, int lane, int col_off = 0, int n_cols = BN,
                                              int part = 0, int parts = 1, const float* sk_src = nullptr,
                                              int sk_n = 0, int64_t sk_stride = 0, bool sk_smem = false) {
      // sk_src / sk_n / sk_stride: split-K partials of sk_n other CTAs for this thread's row (workspace layout of
      // gemm_tcgen05.cu: chunk c, column group g at (c * 8 + g) * 512 floats), added to every chunk read from TMEM
      auto sk_add = [&](uint32_t (&r)[32], int chunk) {
        for (int c = 0; c < sk_n; ++c) {
          const float* w = sk_src + (int64_t)c * sk_stride + chunk * 4096;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            // partials in the workspace were written by other SMs: L2 only; cluster split-K parks them in this CTA's smem
            const float4 v = sk_smem ? *reinterpret_cast<const float4*>(w + g * 512)
                                     : __ldcg(reinterpret_cast<const float4*>(w + g * 512));
            r[4 * g] = __float_as_uint(__uint_as_float(r[4 * g]) + v.x);
            r[4 * g + 1] = __float_as_uint(__uint_as_float(r[4 * g + 1]) + v.y);
            r[4 * g + 2] = __float_as_uint(__uint_as_float(r[4 * g + 2]) + v.z);
            r[4 * g + 3] = __float_as_uint(__uint_as_float(r[4 * g + 3]) + v.w);
          }
        }
      };
      // part / parts: this warp takes the 32-column chunks part, part + parts, ... (epilogue_parts() below says when the
      // second warp set may join: never on the paths that stage through stage_buf)
      const int row = row0 + lane;                       // the accumulator row this thread reads
      if constexpr (EPI == MTS_EPI_ROPE_QK) {
        // q/k columns are rotated in fp32 straight from the accumulators: x1' = x1 cos - x2 sin,
        // x2' = x2 cos + x1 sin with x2 = the column hd/2 further (HF:models/llama/modeling_llama.py:139-168),
        // position = row index inside the sample.  Chunk pairs (c, c + hd/64) of 32 columns hold (x1, x2).
        const int half_chunks = p.rope_hd / 64;           // 1 (hd 64) or 2 (hd 128)
        // shared-prefix layout: rows < rope_prefix are positions themselves, then rope_L own rows per sample
        const int pos = row < p.rope_prefix ? row : p.rope_prefix + (row - p.rope_prefix) % p.rope_L;
        const float* cr = p.rope_cos + (int64_t)pos * (p.rope_hd / 2);
        const float* sr = p.rope_sin + (int64_t)pos * (p.rope_hd / 2);
        __nv_bfloat16* dbase = reinterpret_cast<__nv_bfloat16*>(p.d) + (int64_t)b * p.d_batch_stride;
        const int rr = lane >> 2, cc = (lane & 3) * 8;
        // one 32-column chunk of rotated values -> D
        auto store_chunk = [&](const float (&v)[32], int col0) {
          if (!p.d_is_f32 && p.direct && col0 + 32 <= p.n) {
            // direct path (as in the generic epilogue below): the lane's 32 bf16 columns are 64 contiguous bytes
            if (row < p.m) {
              __nv_bfloat16* dptr = dbase + (int64_t)row * p.ldd + col0;
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = pack_bf16(v[16 * q + 2 * e], v[16 * q + 2 * e + 1]);
                st_global_v8(dptr + 16 * q, w);
              }
            }
            return;
          }
          __syncwarp();
          if (p.d_is_f32) {
            // fp32 output (evaluation parity modes; q / k / v feed the fp32 attention unrounded unless asked):
            // 8 lanes x float4 per row, 4 rows per pass
            const bool rnd = p.round_tf32 != 0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(stage_buf + lane * kEpiPitch + j) =
                  rnd ? make_float4(round_tf32(v[j]), round_tf32(v[j + 1]), round_tf32(v[j + 2]), round_tf32(v[j + 3]))
                      : make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            float* fbase = reinterpret_cast<float*>(p.d) + (int64_t)b * p.d_batch_stride;
            const int rr4 = lane >> 3, cc4 = (lane & 7) * 4;
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int r_g = row0 + ps * 4 + rr4;
              if (col0 + cc4 < p.n && r_g < p.m)
                *reinterpret_cast<float4*>(fbase + (int64_t)r_g * p.ldd + col0 + cc4) =
                    *reinterpret_cast<const float4*>(stage_buf + (ps * 4 + rr4) * kEpiPitch + cc4);
            }
            return;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stage_buf + lane * kEpiPitch + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const int r_g = row0 + ps * 8 + rr;
            if (col0 + cc < p.n && r_g < p.m) {
              const float4 a0 = *reinterpret_cast<const float4*>(stage_buf + (ps * 8 + rr) * kEpiPitch + cc);
              const float4 a1 = *reinterpret_cast<const float4*>(stage_buf + (ps * 8 + rr) * kEpiPitch + cc + 4);
              *reinterpret_cast<uint4*>(dbase + (int64_t)r_g * p.ldd + col0 + cc) =
                  make_uint4(pack_bf16(a0.x, a0.y), pack_bf16(a0.z, a0.w), pack_bf16(a1.x, a1.y),
                             pack_bf16(a1.z, a1.w));
            }
          }
        };
        int pair_idx = 0;                                              // (x1, x2) chunk pairs are dealt out to the warp sets
#pragma unroll 1
        for (int c0 = 0; c0 < BN / 32; c0 += 2 * half_chunks) {        // one head (hd columns) per iteration
#pragma unroll 1
          for (int h = 0; h < half_chunks; ++h, ++pair_idx) {
            if ((pair_idx % parts) != part) continue;                   // warp-uniform
            const int ca = c0 + h, cb = ca + half_chunks;
            uint32_t xa[32], xb[32];
            __syncwarp();
            tmem_ld_32x32(taddr + ca * 32, xa);
            tmem_ld_32x32(taddr + cb * 32, xb);
            tmem_ld_wait();
            if (sk_n > 0) { sk_add(xa, ca); sk_add(xb, cb); }
            const int col_a = n_blk * BN + ca * 32, col_b = n_blk * BN + cb * 32;
            if (col_a >= p.n) continue;                                 // warp-uniform
            float va[32], vb[32];
            if (col_a < p.rope_cols) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 cv = __ldg(reinterpret_cast<const float4*>(cr + h * 32 + j));
                const float4 sv = __ldg(reinterpret_cast<const float4*>(sr + h * 32 + j));
                const float cs[4] = {cv.x, cv.y, cv.z, cv.w}, sn[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float x1 = __uint_as_float(xa[j + q]) * p.alpha, x2 = __uint_as_float(xb[j + q]) * p.alpha;
                  va[j + q] = x1 * cs[q] - x2 * sn[q];
                  vb[j + q] = x2 * cs[q] + x1 * sn[q];
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                va[j] = __uint_as_float(xa[j]) * p.alpha;
                vb[j] = __uint_as_float(xb[j]) * p.alpha;
              }
            }
            store_chunk(va, col_a);
            store_chunk(vb, col_b);
          }
        }
        return;
      }
      const float bias_m = (p.bias_axis == 2 && row < p.m) ? p.bias[row] : 0.0f;

      const int kChunks = (EPI == MTS_EPI_SWIGLU) ? BN / 64 : n_cols / 32;
      const int n_store = (EPI == MTS_EPI_SWIGLU) ? p.n / 2 : p.n;
#pragma unroll 1
      for (int ci = part; ci < kChunks; ci += parts) {
        float v[32];
        int col0;  // first output column of this chunk
        __syncwarp();  // tcgen05.ld is .sync.aligned; also orders the previous chunk's smem reads
        if constexpr (EPI == MTS_EPI_SWIGLU) {
          // columns [0,BN/2) of the tile are gate, [BN/2,BN) the matching up projections
          uint32_t g[32], u[32];
          tmem_ld_32x32(taddr + ci * 32, g);
          tmem_ld_32x32(taddr + BN / 2 + ci * 32, u);
          tmem_ld_wait();
          if (sk_n > 0) { sk_add(g, ci); sk_add(u, BN / 64 + ci); }
          col0 = n_blk * (BN / 2) + ci * 32;
          if (p.aux != nullptr && row < p.m && n_blk * BN + ci * 32 < p.n) {
            // pre-activations for the backward: 64 contiguous bytes per thread and half (16-byte stores)
            __nv_bfloat16* ag = p.aux + (int64_t)b * p.m * p.ld_aux + (int64_t)row * p.ld_aux + n_blk * BN + ci * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              *reinterpret_cast<uint4*>(ag + j) = make_uint4(
                  pack_bf16(__uint_as_float(g[j]) * p.alpha, __uint_as_float(g[j + 1]) * p.alpha),
                  pack_bf16(__uint_as_float(g[j + 2]) * p.alpha, __uint_as_float(g[j + 3]) * p.alpha),
                  pack_bf16(__uint_as_float(g[j + 4]) * p.alpha, __uint_as_float(g[j + 5]) * p.alpha),
                  pack_bf16(__uint_as_float(g[j + 6]) * p.alpha, __uint_as_float(g[j + 7]) * p.alpha));
              *reinterpret_cast<uint4*>(ag + BN / 2 + j) = make_uint4(
                  pack_bf16(__uint_as_float(u[j]) * p.alpha, __uint_as_float(u[j + 1]) * p.alpha),
                  pack_bf16(__uint_as_float(u[j + 2]) * p.alpha, __uint_as_float(u[j + 3]) * p.alpha),
                  pack_bf16(__uint_as_float(u[j + 4]) * p.alpha, __uint_as_float(u[j + 5]) * p.alpha),
                  pack_bf16(__uint_as_float(u[j + 6]) * p.alpha, __uint_as_float(u[j + 7]) * p.alpha));
            }
          }
          if (p.precise) {        // evaluation parity modes: library-accurate expf (warp-uniform branch, hoisted)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = silu_precise_f(__uint_as_float(g[j]) * p.alpha) * (__uint_as_float(u[j]) * p.alpha);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              // the activation is computed from the bf16-rounded pre-activations when they are kept, so that the
              // saved tensors reproduce exactly what the forward used
              float gv = __uint_as_float(g[j]) * p.alpha, uv = __uint_as_float(u[j]) * p.alpha;
              if (p.aux != nullptr) { gv = __bfloat162float(__float2bfloat16_rn(gv)); uv = __bfloat162float(__float2bfloat16_rn(uv)); }
              v[j] = silu_f(gv) * uv;
            }
          }
        } else {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ci * 32, r);
          tmem_ld_wait();
          if (sk_n > 0) sk_add(r, ci);
          col0 = n_blk * BN + col_off + ci * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha + bias_m;
          if (p.bias_axis == 1 && col0 < p.n) {
            if (p.bias_vec && col0 + 32 <= p.n) {  // warp-uniform 16-byte loads (broadcast)
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n) v[j] += __ldg(p.bias + col0 + j);
            }
          }
          if constexpr (EPI == MTS_EPI_GELU_NEW) {
            if (p.precise) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_new_precise_f(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_new_f(v[j]);
            }
          }
        }
        if (col0 >= n_store) continue;  // warp-uniform
        if constexpr (EPI == MTS_EPI_RESID_ADD) {
          if (p.drop_thresh != 0) {
            const uint64_t base = ((uint64_t)b * p.m + (uint64_t)row) * (uint64_t)p.n + (uint64_t)col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = dropout_keep(p.drop_seed, base + j, p.drop_thresh) ? v[j] * p.drop_scale : 0.0f;
          }
        }
        if constexpr (EPI != MTS_EPI_RESID_ADD) {
          if (p.round_tf32) {   // fp32 result that is the operand of a kind::tf32 GEMM: round to nearest, not truncate
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
          }
        }

        if (EPI == MTS_EPI_STORE && p.d_transposed) {
          // element (row, col) -> d[col*ldd + row]: lanes (= rows) are contiguous in memory
          if (row < p.m) {
            if (p.d_is_f32) {
              float* dptr = reinterpret_cast<float*>(p.d) + (int64_t)b * p.d_batch_stride + row;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n) dptr[(int64_t)(col0 + j) * p.ldd] = v[j];
            } else {
              __nv_bfloat16* dptr =
                  reinterpret_cast<__nv_bfloat16*>(p.d) + (int64_t)b * p.d_batch_stride + row;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n) dptr[(int64_t)(col0 + j) * p.ldd] = __float2bfloat16_rn(v[j]);
            }
          }
          continue;
        }

        if (p.direct && col0 + 32 <= n_store) {
          // direct path: this lane's 32 columns of row `row` are contiguous in memory — whole 32-byte sectors per access,
          // no shared-memory round trip, no warp synchronisation (the exposed epilogue of a one-tile-per-CTA launch is a
          // serial chain on four warps: every dependent step removed is time off the launch)
          if (row < p.m) {
            const int64_t off = (int64_t)b * p.d_batch_stride + (int64_t)row * p.ldd + col0;
            if (p.d_is_f32) {
              float* dptr = reinterpret_cast<float*>(p.d) + off;
              if constexpr (EPI == MTS_EPI_RESID_ADD) {
                const float* cptr = p.c + off;
                uint32_t cw[4][8];
#pragma unroll
                for (int q = 0; q < 4; ++q) ld_global_v8(cptr + 8 * q, cw[q]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) cw[q][e] = __float_as_uint(__uint_as_float(cw[q][e]) + v[8 * q + e]);
                  st_global_v8(dptr + 8 * q, cw[q]);
                }
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint32_t w[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(v[8 * q + e]);
                  st_global_v8(dptr + 8 * q, w);
                }
              }
            } else {
              __nv_bfloat16* dptr = reinterpret_cast<__nv_bfloat16*>(p.d) + off;
              if constexpr (EPI == MTS_EPI_RESID_ADD) {   // bf16 accumulate (LoRA side GEMMs): D = bf16(D + v)
                uint32_t ow[2][8];
                ld_global_v8(dptr, ow[0]);
                ld_global_v8(dptr + 16, ow[1]);
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                  for (int e = 0; e < 8; ++e) { v[16 * q + 2 * e] += bf16_lo(ow[q][e]); v[16 * q + 2 * e + 1] += bf16_hi(ow[q][e]); }
              }
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = pack_bf16(v[16 * q + 2 * e], v[16 * q + 2 * e + 1]);
                st_global_v8(dptr + 16 * q, w);
              }
            }
          }
          continue;
        }

        // stage the 32x32 chunk: thread `lane` owns row `lane` (pitch 36 floats: 16-byte aligned
        // rows, conflict-free for both the 128-bit writes here and the 128-bit reads below)
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage_buf + lane * kEpiPitch + j) =
              make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();

        if (!p.vec_ok) {
          // rows of D are not 16-byte aligned (odd n / ldd): scalar, still row-contiguous per lane
#pragma unroll 1
          for (int rr = 0; rr < 32; ++rr) {
            const int r_g = row0 + rr;
            const int col = col0 + lane;
            if (r_g < p.m && col < n_store) {
              const int64_t off = (int64_t)b * p.d_batch_stride + (int64_t)r_g * p.ldd + col;
              float val = stage_buf[rr * kEpiPitch + lane];
              if (p.d_is_f32) {
                if constexpr (EPI == MTS_EPI_RESID_ADD) val += p.c[off];
                reinterpret_cast<float*>(p.d)[off] = val;
              } else {
                if constexpr (EPI == MTS_EPI_RESID_ADD)
                  val += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.d)[off]);
                reinterpret_cast<__nv_bfloat16*>(p.d)[off] = __float2bfloat16_rn(val);
              }
            }
          }
        } else if (p.d_is_f32) {
          // 8 lanes x float4 per row, 4 rows per pass
          const int rr = lane >> 3, cc = (lane & 7) * 4;
          const int64_t boff = (int64_t)b * p.d_batch_stride + col0 + cc;
          float* dbase = reinterpret_cast<float*>(p.d) + boff;
          const bool col_ok = col0 + cc < n_store;
          if constexpr (EPI == MTS_EPI_RESID_ADD) {
            const float* cbase = p.c + boff;  // residual source (== D for the in-place form)
            float4 o[8];
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int r_g = row0 + ps * 4 + rr;
              if (col_ok && r_g < p.m) o[ps] = *reinterpret_cast<const float4*>(cbase + (int64_t)r_g * p.ldd);
            }
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int r_g = row0 + ps * 4 + rr;
              if (col_ok && r_g < p.m) {
                const float4 a = *reinterpret_cast<const float4*>(stage_buf + (ps * 4 + rr) * kEpiPitch + cc);
                o[ps].x += a.x; o[ps].y += a.y; o[ps].z += a.z; o[ps].w += a.w;
                *reinterpret_cast<float4*>(dbase + (int64_t)r_g * p.ldd) = o[ps];
              }
            }
          } else {
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int r_g = row0 + ps * 4 + rr;
              if (col_ok && r_g < p.m)
                *reinterpret_cast<float4*>(dbase + (int64_t)r_g * p.ldd) =
                    *reinterpret_cast<const float4*>(stage_buf + (ps * 4 + rr) * kEpiPitch + cc);
            }
          }
        } else {
          // bf16: 4 lanes x 8 columns (16 bytes) per row, 8 rows per pass
          const int rr = lane >> 2, cc = (lane & 3) * 8;
          __nv_bfloat16* dbase =
              reinterpret_cast<__nv_bfloat16*>(p.d) + (int64_t)b * p.d_batch_stride + col0 + cc;
          const bool col_ok = col0 + cc < n_store;
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const int r_g = row0 + ps * 8 + rr;
            if (col_ok && r_g < p.m) {
              float4 a0 = *reinterpret_cast<const float4*>(stage_buf + (ps * 8 + rr) * kEpiPitch + cc);
              float4 a1 = *reinterpret_cast<const float4*>(stage_buf + (ps * 8 + rr) * kEpiPitch + cc + 4);
              if constexpr (EPI == MTS_EPI_RESID_ADD) {   // bf16 accumulate (LoRA side GEMMs): D = bf16(D + v)
                const uint4 old = *reinterpret_cast<const uint4*>(dbase + (int64_t)r_g * p.ldd);
                a0.x += bf16_lo(old.x); a0.y += bf16_hi(old.x); a0.z += bf16_lo(old.y); a0.w += bf16_hi(old.y);
                a1.x += bf16_lo(old.z); a1.y += bf16_hi(old.z); a1.z += bf16_lo(old.w); a1.w += bf16_hi(old.w);
              }
              *reinterpret_cast<uint4*>(dbase + (int64_t)r_g * p.ldd) =
                  make_uint4(pack_bf16(a0.x, a0.y), pack_bf16(a0.z, a0.w), pack_bf16(a1.x, a1.y),
                             pack_bf16(a1.z, a1.w));
            }
          }
        }
      }
}

}  // namespace mts
