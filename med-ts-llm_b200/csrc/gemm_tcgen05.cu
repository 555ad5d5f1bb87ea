// gemm_tcgen05.cu — the one GEMM of the hot path:  D[b] = epi(alpha * A[b] * B[b]^T + bias)
//
// Both operands are K-major bf16 (activations [m,k]; nn.Linear weights [n,k]), which is what the
// reference's Linear/Conv1D/einsum call sites reduce to (see include/mts_b200.h for the list).
//
// Structure (B200, one CTA per SM, persistent over output tiles):
//   warp 0 / lane 0 : TMA producer  — cp.async.bulk.tensor 128x64 (A) and BNx64 (B) bf16 boxes
//                     into a kStages-deep ring of 128B-swizzled smem tiles, signalled by mbarriers
//   warp 1 / lane 0 : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 x4 per
//                     stage; fp32 accumulators in TMEM, two accumulator stages (2*BN columns) so
//                     the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2..5      : epilogue      — tcgen05.ld 32x32b.x32 (thread = row, 32 columns), fused
//                     bias / residual-add / gelu_new / silu*up, vectorised global stores
//
// Algorithmic work per launch: 2*m*n*k*batch FLOP (roofline: tensor pipe).
#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;
constexpr int kNumEpiThreads = 128;
constexpr int kEpiPitch = 36;  // floats per staged row (144 B: 16-byte aligned, bank-conflict free)

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // leave ~3 KB for barriers + alignment slack out of the 227 KB budget
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;  // 128, 256 or 512: all powers of two >= 32
  static constexpr int kBarrierBytes = 256;
  static constexpr int kEpiStageBytes = 4 * 32 * kEpiPitch * 4;  // per-epilogue-warp 32x32 fp32 tiles
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + kEpiStageBytes + 1024;
};

struct GemmParams {
  void* d;
  const float* bias;
  int64_t ldd, d_batch_stride;
  int m, n, k, batch;
  int a_batched, b_batched;  // 0: operand shared by all batches
  int bias_axis, d_transposed, d_is_f32, bias_vec, vec_ok;
  const float* c;  // RESID_ADD: residual source, same layout as d
  float alpha;
};

__device__ __forceinline__ float gelu_new_f(float x) {
  // HF:activations.py:65-66  0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));  // |err| ~ 2^-11: below the bf16 output rounding
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

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

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_nt_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStages * Cfg::kABytes;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout (8 bytes each): full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  const uint32_t stage_base = bar_base + Cfg::kBarrierBytes;  // 16-byte aligned
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (p.m + kBlockM - 1) / kBlockM;
  const int n_blocks = (p.n + BN - 1) / BN;
  const int k_blocks = (p.k + kBlockK - 1) / kBlockK;
  const int tiles_per_batch = m_blocks * n_blocks;
  const int num_tiles = tiles_per_batch * p.batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kNumEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_batch;
        const int t = tile - b * tiles_per_batch;
        int m_blk, n_blk;
        tile_coords(t, m_blocks, n_blocks, m_blk, n_blk);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_3d(smem_a + stage * Cfg::kABytes, &tmap_a, full_bar(stage), kb * kBlockK,
                      m_blk * kBlockM, p.a_batched ? b : 0, kEvictNormal);
          tma_load_3d(smem_b + stage * Cfg::kBBytes, &tmap_b, full_bar(stage), kb * kBlockK,
                      n_blk * BN, p.b_batched ? b : 0, kEvictNormal);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 200 + acc);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_a + stage * Cfg::kABytes);
          const uint64_t bdesc = umma_desc_sw128(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzled row: +2 in (addr >> 4)
            umma_bf16(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5
    // TMEM -> registers (thread = accumulator row, 32 columns) -> per-warp padded smem tile ->
    // coalesced 16-byte global accesses (a quarter-warp covers one contiguous 128-byte row piece).
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    float* stage_buf = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw))) +
                       quarter * (32 * kEpiPitch);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_batch;
      const int t = tile - b * tiles_per_batch;
      int m_blk, n_blk;
      tile_coords(t, m_blocks, n_blocks, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();

      const int row0 = m_blk * kBlockM + quarter * 32;  // first row of this warp's slab
      const int row = row0 + lane;                       // the accumulator row this thread reads
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
      const float bias_m = (p.bias_axis == 2 && row < p.m) ? p.bias[row] : 0.0f;

      constexpr int kChunks = (EPI == MTS_EPI_SWIGLU) ? BN / 64 : BN / 32;
      const int n_store = (EPI == MTS_EPI_SWIGLU) ? p.n / 2 : p.n;
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        float v[32];
        int col0;  // first output column of this chunk
        __syncwarp();  // tcgen05.ld is .sync.aligned; also orders the previous chunk's smem reads
        if constexpr (EPI == MTS_EPI_SWIGLU) {
          // columns [0,BN/2) of the tile are gate, [BN/2,BN) the matching up projections
          uint32_t g[32], u[32];
          tmem_ld_32x32(taddr + ci * 32, g);
          tmem_ld_32x32(taddr + BN / 2 + ci * 32, u);
          tmem_ld_wait();
          col0 = n_blk * (BN / 2) + ci * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = silu_f(__uint_as_float(g[j]) * p.alpha) * (__uint_as_float(u[j]) * p.alpha);
        } else {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ci * 32, r);
          tmem_ld_wait();
          col0 = n_blk * BN + ci * 32;
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
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_new_f(v[j]);
          }
        }
        if (col0 >= n_store) continue;  // warp-uniform

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
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       int num_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_nt_kernel<BN, EPI>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gemm smem)", e);
    attr_done = true;
  }
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  count_launch();
  return check_launch("gemm_bf16_nt_kernel");
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                        int num_tiles, cudaStream_t stream) {
  switch (epi) {
    case MTS_EPI_STORE: return launch_gemm<BN, MTS_EPI_STORE>(ta, tb, p, num_tiles, stream);
    case MTS_EPI_RESID_ADD: return launch_gemm<BN, MTS_EPI_RESID_ADD>(ta, tb, p, num_tiles, stream);
    case MTS_EPI_GELU_NEW: return launch_gemm<BN, MTS_EPI_GELU_NEW>(ta, tb, p, num_tiles, stream);
    default: return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: unknown epilogue");
  }
}

static int pick_block_n(int m, int n, int batch) {
  // minimise waves x (measured relative time of one tile).  A 128x128 tile is shared-memory-bandwidth
  // bound (A and B are each re-read per MMA): it costs ~0.77 of a 128x256 tile, not 0.5; a 128x64
  // tile ~0.6 (B200 measurements, tools/bench_gemm.py).  Narrow tiles only win on small problems.
  const int mb = (m + kBlockM - 1) / kBlockM;
  int best = 64;
  long best_cost = -1;
  const int cands[3] = {256, 128, 64};
  const int tile_cost[3] = {280, 209, 160};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long tiles = (long)mb * ((n + bn - 1) / bn) * batch;
    const long waves = (tiles + num_sms() - 1) / num_sms();
    const long cost = waves * tile_cost[i];
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace mts

using namespace mts;

extern "C" int mts_gemm(const mts_gemm_args* a, mts_stream_t stream_) {
  if (!a) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: null args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a->m <= 0 || a->n <= 0 || a->k <= 0 || a->batch <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: m, n, k, batch must be positive");
  if (!a->a || !a->b || !a->d) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: null operand");
  if ((a->lda % 8) || (a->ldb % 8) || a->lda < a->k || a->ldb < a->k)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: lda/ldb must be >= k and multiples of 8");
  if ((reinterpret_cast<uintptr_t>(a->a) & 15) || (reinterpret_cast<uintptr_t>(a->b) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: a and b must be 16-byte aligned");
  if ((a->a_batch_stride % 8) || (a->b_batch_stride % 8))
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: batch strides must be multiples of 8");
  if (a->bias_axis != MTS_BIAS_NONE && !a->bias)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: bias_axis set but bias is null");
  const bool f32 = a->d_dtype == MTS_F32;
  if (a->d_dtype != MTS_F32 && a->d_dtype != MTS_BF16)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: bad d_dtype");

  int bn = a->block_n;
  const int n_store = a->epilogue == MTS_EPI_SWIGLU ? a->n / 2 : a->n;
  switch (a->epilogue) {
    case MTS_EPI_STORE:
      break;
    case MTS_EPI_RESID_ADD:
      if (a->d_transposed || (!f32 && a->c))
        return set_error(MTS_ERR_INVALID_ARG,
                         "mts_gemm: RESID_ADD needs a non-transposed D (fp32 with optional C, or bf16 in place)");
      break;
    case MTS_EPI_GELU_NEW:
      if (f32 || a->d_transposed)
        return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: GELU_NEW needs bf16 non-transposed D");
      break;
    case MTS_EPI_SWIGLU:
      if (f32 || a->d_transposed || a->bias_axis != MTS_BIAS_NONE || (a->n % 256) ||
          (bn != 0 && bn != 256))
        return set_error(MTS_ERR_INVALID_ARG,
                         "mts_gemm: SWIGLU needs bf16 D, no bias, n % 256 == 0 (packed gate/up), "
                         "block_n 256");
      bn = 256;
      break;
    default:
      return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: unknown epilogue");
  }
  int vec_ok = 1;
  if (!a->d_transposed) {
    const int vec = f32 ? 4 : 8;
    if ((n_store % vec) || (a->ldd % vec) || (a->d_batch_stride % vec) ||
        (reinterpret_cast<uintptr_t>(a->d) & 15) || (a->c && (reinterpret_cast<uintptr_t>(a->c) & 15)))
      vec_ok = 0;  // scalar epilogue stores
    if (a->ldd < n_store) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: ldd < n");
  }
  if (a->c && a->epilogue != MTS_EPI_RESID_ADD)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: c is only used by MTS_EPI_RESID_ADD");
  if (bn == 0) bn = pick_block_n(a->m, a->n, a->batch);
  if (bn != 64 && bn != 128 && bn != 256)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: block_n must be 0, 64, 128 or 256");

  const int a_batched = (a->batch > 1 && a->a_batch_stride != 0) ? 1 : 0;
  const int b_batched = (a->batch > 1 && a->b_batch_stride != 0) ? 1 : 0;
  CUtensorMap ta, tb;
  int rc = get_tmap_bf16_3d(&ta, a->a, a->k, a->m, a_batched ? a->batch : 1, a->lda,
                            a_batched ? a->a_batch_stride : (int64_t)a->m * a->lda, kBlockK,
                            kBlockM);
  if (rc) return rc;
  rc = get_tmap_bf16_3d(&tb, a->b, a->k, a->n, b_batched ? a->batch : 1, a->ldb,
                        b_batched ? a->b_batch_stride : (int64_t)a->n * a->ldb, kBlockK, bn);
  if (rc) return rc;

  GemmParams p;
  p.d = a->d;
  p.bias = a->bias;
  p.ldd = a->ldd;
  p.d_batch_stride = a->d_batch_stride;
  p.m = a->m; p.n = a->n; p.k = a->k; p.batch = a->batch;
  p.a_batched = a_batched; p.b_batched = b_batched;
  p.bias_axis = a->bias_axis;
  p.d_transposed = a->d_transposed;
  p.d_is_f32 = f32 ? 1 : 0;
  p.vec_ok = vec_ok;
  p.c = a->c ? a->c : static_cast<const float*>(a->d);
  p.bias_vec = (a->bias && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0) ? 1 : 0;
  p.alpha = a->alpha;

  const int mb = (a->m + kBlockM - 1) / kBlockM;
  const int nb = (a->n + bn - 1) / bn;
  const long tiles_l = (long)mb * nb * a->batch;
  if (tiles_l > 0x7fffffffL) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: too many tiles");
  const int tiles = (int)tiles_l;

  if (a->epilogue == MTS_EPI_SWIGLU) return launch_gemm<256, MTS_EPI_SWIGLU>(ta, tb, p, tiles, stream);
  switch (bn) {
    case 256: return dispatch_epi<256>(a->epilogue, ta, tb, p, tiles, stream);
    case 128: return dispatch_epi<128>(a->epilogue, ta, tb, p, tiles, stream);
    default: return dispatch_epi<64>(a->epilogue, ta, tb, p, tiles, stream);
  }
}
