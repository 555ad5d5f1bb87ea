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

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // leave ~3 KB for barriers + alignment slack out of the 227 KB budget
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;  // 128, 256 or 512: all powers of two >= 32
  static constexpr int kBarrierBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;
};

struct GemmParams {
  void* d;
  const float* bias;
  int64_t ldd, d_batch_stride;
  int m, n, k, batch;
  int a_batched, b_batched;  // 0: operand shared by all batches
  int bias_axis, d_transposed, d_is_f32;
  float alpha;
};

__device__ __forceinline__ float gelu_new_f(float x) {
  // HF:activations.py:65-66  0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

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
        const int n_blk = t / m_blocks;  // m fastest: CTAs running together share the B tile
        const int m_blk = t - n_blk * m_blocks;
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
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_batch;
      const int t = tile - b * tiles_per_batch;
      const int n_blk = t / m_blocks;
      const int m_blk = t - n_blk * m_blocks;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();

      const int row = m_blk * kBlockM + row_in_tile;
      const bool row_ok = row < p.m;
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
      const float bias_m = (p.bias_axis == 2 && row_ok) ? p.bias[row] : 0.0f;

      if constexpr (EPI == MTS_EPI_SWIGLU) {
        // columns [0,BN/2) of the tile are gate, [BN/2,BN) the matching up projections
        constexpr int kHalf = BN / 2;
        __nv_bfloat16* dptr = reinterpret_cast<__nv_bfloat16*>(p.d) + (int64_t)b * p.d_batch_stride +
                              (int64_t)row * p.ldd;
        const int n_out = p.n / 2;
#pragma unroll 1
        for (int c = 0; c < kHalf; c += 32) {
          uint32_t g[32], u[32];
          __syncwarp();
          tmem_ld_32x32(taddr + c, g);
          tmem_ld_32x32(taddr + kHalf + c, u);
          tmem_ld_wait();
          const int col0 = n_blk * kHalf + c;
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < n_out) {
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float g0 = __uint_as_float(g[j + 2 * q]) * p.alpha;
                  const float g1 = __uint_as_float(g[j + 2 * q + 1]) * p.alpha;
                  const float u0 = __uint_as_float(u[j + 2 * q]) * p.alpha;
                  const float u1 = __uint_as_float(u[j + 2 * q + 1]) * p.alpha;
                  o[q] = pack_bf16(silu_f(g0) * u0, silu_f(g1) * u1);
                }
                *reinterpret_cast<uint4*>(dptr + col0 + j) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t r[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the guarded stores
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c;
          if (row_ok && col0 < p.n) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha + bias_m;
          if (p.bias_axis == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.n) v[j] += __ldg(p.bias + col0 + j);
          }
          if constexpr (EPI == MTS_EPI_GELU_NEW) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_new_f(v[j]);
          }
          if constexpr (EPI == MTS_EPI_RESID_ADD) {
            float* dptr = reinterpret_cast<float*>(p.d) + (int64_t)b * p.d_batch_stride +
                          (int64_t)row * p.ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.n) {  // n % 4 == 0 enforced on the host
                float4 o = *reinterpret_cast<const float4*>(dptr + j);
                o.x += v[j]; o.y += v[j + 1]; o.z += v[j + 2]; o.w += v[j + 3];
                *reinterpret_cast<float4*>(dptr + j) = o;
              }
            }
          } else if (p.d_transposed) {
            // element (row, col) -> d[col*ldd + row]; small matrices only (head staging)
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
          } else if (p.d_is_f32) {
            float* dptr = reinterpret_cast<float*>(p.d) + (int64_t)b * p.d_batch_stride +
                          (int64_t)row * p.ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.n)
                *reinterpret_cast<float4*>(dptr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            __nv_bfloat16* dptr = reinterpret_cast<__nv_bfloat16*>(p.d) +
                                  (int64_t)b * p.d_batch_stride + (int64_t)row * p.ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < p.n) {  // n % 8 == 0 enforced on the host
                *reinterpret_cast<uint4*>(dptr + j) =
                    make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                               pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
              }
            }
          }
          }  // row_ok && col0 < n
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
  // minimise (waves x tile width); prefer the wider tile on ties (fewer B re-reads of A)
  const int mb = (m + kBlockM - 1) / kBlockM;
  int best = 64;
  long best_cost = -1;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long tiles = (long)mb * ((n + bn - 1) / bn) * batch;
    const long waves = (tiles + num_sms() - 1) / num_sms();
    const long cost = waves * (bn + 24);  // +24: per-tile pipeline fill/drain, in "columns"
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
      if (!f32 || a->d_transposed)
        return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: RESID_ADD needs fp32 non-transposed D");
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
  if (!a->d_transposed) {
    const int vec = f32 ? 4 : 8;
    if ((n_store % vec) || (a->ldd % vec) || (a->d_batch_stride % vec) ||
        (reinterpret_cast<uintptr_t>(a->d) & 15))
      return set_error(MTS_ERR_INVALID_ARG,
                       "mts_gemm: D needs 16-byte aligned rows (n, ldd, batch stride multiples of "
                       "8 for bf16 / 4 for fp32)");
  }
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
