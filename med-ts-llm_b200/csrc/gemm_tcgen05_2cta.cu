// gemm_tcgen05_2cta.cu — CTA-pair variant of the NT GEMM: D = epi(alpha * A * B^T + bias)
//
// A cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  CTA r stages A rows [r*128, +128) and
// the r-th HALF of the B tile (128 of its 256 rows) per k-block, so a pair moves 64 KB instead of the
// 96 KB two independent 128x256 CTAs would (-33 % L2->SM and shared-memory fill traffic).  Only the
// leader CTA issues `tcgen05.mma.cta_group::2` (M = 256, N = 256, K = 16); the tensor cores of both
// SMs read both halves of B.  Each CTA keeps its own 128 x 256 fp32 accumulator (double buffered) in
// its own TMEM and runs the same fused epilogue as the 1-CTA kernel on its own rows.
//
// Barriers (same smem offsets in both CTAs):
//   full[s]   leader only; 2 arrivals (leader's arrive.expect_tx of 64 KB + peer's remote arrive);
//             all four TMA loads of a stage credit their bytes to the leader's barrier
//   empty[s]  per CTA; released by the leader's multicast tcgen05.commit
//   tfull[a]  per CTA; multicast commit after the tile's last k-block
//   tempty[a] leader only; 512 arrivals (both CTAs' epilogue threads, the peer's remotely)
#include "gemm_common.cuh"

namespace mts {

constexpr int kPairBN = 256;
constexpr int kPairStages = 6;
constexpr int kPairABytes = kBlockM * kBlockK * 2;          // 16 KB: this CTA's 128 rows of A
constexpr int kPairBBytes = (kPairBN / 2) * kBlockK * 2;    // 16 KB: this CTA's half of B
constexpr int kPairStageBytes = kPairABytes + kPairBBytes;
constexpr int kPairBarrierBytes = 256;
constexpr int kPairSmemBytes =
    kPairStages * kPairStageBytes + kPairBarrierBytes + 4 * 32 * kEpiPitch * 4 + 1024;

// Work units of the CTA-pair kernel.  Units [0, full) are whole 256x256 tiles.  When the last wave would
// be mostly empty (e.g. 384 tiles on 74 pairs = 5.19 waves: o-proj / down-proj / every dgrad with N = 4096),
// its `tail` tiles are split into `split` column slices of 256/split columns each, so the tail wave costs a
// fraction of a full tile instead of a whole one.
struct PairSchedule {
  int full, split, total;   // total = full + tail * split
};
__host__ __device__ inline PairSchedule pair_schedule(int num_tiles, int num_pairs, bool allow_split) {
  PairSchedule s;
  const int tail = num_tiles % num_pairs;
  s.split = 1;
  if (allow_split && num_tiles > num_pairs && tail > 0) {
    if (tail * 4 <= num_pairs) s.split = 4;
    else if (tail * 2 <= num_pairs) s.split = 2;
  }
  s.full = s.split > 1 ? num_tiles - tail : num_tiles;
  s.total = s.full + (num_tiles - s.full) * s.split;
  return s;
}
struct PairUnit {
  int tile, col_off, n_cols;
};
__device__ __forceinline__ PairUnit pair_unit(int u, const PairSchedule& s) {
  PairUnit r;
  if (u < s.full) { r.tile = u; r.col_off = 0; r.n_cols = kPairBN; }
  else {
    const int v = u - s.full;
    r.tile = s.full + v / s.split;
    r.n_cols = kPairBN / s.split;
    r.col_off = (v % s.split) * r.n_cols;
  }
  return r;
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_nt_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  constexpr int BN = kPairBN;
  constexpr int kStages = kPairStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStages * kPairABytes;
  const uint32_t bar_base = smem_base + kStages * kPairStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  const uint32_t stage_base = bar_base + kPairBarrierBytes;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int m_blocks = (p.m + 2 * kBlockM - 1) / (2 * kBlockM);   // 256-row tiles
  const int n_blocks = (p.n + BN - 1) / BN;
  const int k_blocks = (p.k + kBlockK - 1) / kBlockK;
  const int tiles_per_batch = m_blocks * n_blocks;
  const int num_tiles = tiles_per_batch * p.batch;
  const PairSchedule sched = pair_schedule(num_tiles, num_pairs, EPI != MTS_EPI_SWIGLU && EPI != MTS_EPI_ROPE_QK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * kNumEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  cluster_sync_all();                        // barriers of both CTAs exist before any remote arrive
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();       // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = pair; unit < sched.total; unit += num_pairs) {
        const PairUnit pu = pair_unit(unit, sched);
        const int b = pu.tile / tiles_per_batch;
        const int t = pu.tile - b * tiles_per_batch;
        int m_blk, n_blk;
        tile_coords(t, m_blocks, n_blocks, m_blk, n_blk);
        const int row_a = m_blk * 2 * kBlockM + (int)rank * kBlockM;
        // this CTA's half of the unit's B columns (the 128-row box over-reads for narrow units: harmless)
        const int row_b = n_blk * BN + pu.col_off + (int)rank * (pu.n_cols / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * kPairStageBytes);
          else           mbar_arrive_cluster(full_bar(stage), 0);
          tma_load_3d_2sm(smem_a + stage * kPairABytes, &tmap_a, full_bar(stage), kb * kBlockK, row_a,
                          p.a_batched ? b : 0, kEvictNormal);
          tma_load_3d_2sm(smem_b + stage * kPairBBytes, &tmap_b, full_bar(stage), kb * kBlockK, row_b,
                          p.b_batched ? b : 0, kEvictNormal);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = pair; unit < sched.total; unit += num_pairs, ++it) {
        const uint32_t idesc = umma_idesc_bf16(2 * kBlockM, (uint32_t)pair_unit(unit, sched).n_cols);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 200 + acc);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_a + stage * kPairABytes);
          const uint64_t bdesc = umma_desc_sw128(smem_b + stage * kPairBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16_2cta(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta(empty_bar(stage), 0x3);   // both CTAs may refill this stage
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2cta(tfull_bar(acc), 0x3);       // both CTAs' epilogues may read their accumulators
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9 (both CTAs)
    const int quarter = warp & 3;
    const int eset = (warp - 2) >> 2;                       // 0: warps 2..5 (own the staging buffers), 1: warps 6..9
    const int parts = epilogue_parts<EPI>(p);
    float* stage_buf = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw))) +
                       quarter * (32 * kEpiPitch);
    int it = 0;
    for (int unit = pair; unit < sched.total; unit += num_pairs, ++it) {
      const PairUnit pu = pair_unit(unit, sched);
      const int b = pu.tile / tiles_per_batch;
      const int t = pu.tile - b * tiles_per_batch;
      int m_blk, n_blk;
      tile_coords(t, m_blocks, n_blocks, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();
      if (eset < parts)
        epilogue_tile<BN, EPI>(p, b, m_blk * 2 * kBlockM + (int)rank * kBlockM + quarter * 32, n_blk,
                               tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN),
                               stage_buf, lane, pu.col_off, pu.n_cols, eset, parts);
      tc_fence_before();
      if (rank == 0) mbar_arrive(tempty_bar(acc));
      else           mbar_arrive_cluster(tempty_bar(acc), 0);
    }
  }

  __syncwarp();           // reconverge the single-lane roles: barrier.cluster is warp-aligned
  tc_fence_before();
  cluster_sync_all();     // no CTA may exit (or free TMEM) while its partner can still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

template <int EPI>
static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_tiles,
                       cudaStream_t stream) {
  auto kern = gemm_bf16_nt_2cta_kernel<EPI>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gemm 2cta smem)", e);
    attr_done = true;
  }
  const int pairs = num_sms() / 2;
  const int grid = 2 * (num_tiles < pairs ? num_tiles : pairs);   // (splitting only applies when tiles > pairs)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kPairSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  count_launch();
  if (e != cudaSuccess) return set_cuda_error("cudaLaunchKernelEx(gemm_bf16_nt_2cta_kernel)", e);
  return check_launch("gemm_bf16_nt_2cta_kernel");
}

// Called by mts_gemm when the CTA-pair kernel applies (block_n 256, non-transposed D).
int launch_gemm_2cta(int epilogue, const mts_gemm_args* a, const GemmParams& p, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = get_tmap_bf16_3d(&ta, a->a, a->k, a->m, p.a_batched ? a->batch : 1, a->lda,
                            p.a_batched ? a->a_batch_stride : (int64_t)a->m * a->lda, kBlockK, kBlockM);
  if (rc) return rc;
  rc = get_tmap_bf16_3d(&tb, a->b, a->k, a->n, p.b_batched ? a->batch : 1, a->ldb,
                        p.b_batched ? a->b_batch_stride : (int64_t)a->n * a->ldb, kBlockK, kPairBN / 2);
  if (rc) return rc;
  const long tiles_l = (long)((a->m + 255) / 256) * ((a->n + kPairBN - 1) / kPairBN) * a->batch;
  if (tiles_l > 0x7fffffffL) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: too many tiles");
  const int tiles = (int)tiles_l;
  switch (epilogue) {
    case MTS_EPI_STORE: return launch_pair<MTS_EPI_STORE>(ta, tb, p, tiles, stream);
    case MTS_EPI_RESID_ADD: return launch_pair<MTS_EPI_RESID_ADD>(ta, tb, p, tiles, stream);
    case MTS_EPI_GELU_NEW: return launch_pair<MTS_EPI_GELU_NEW>(ta, tb, p, tiles, stream);
    case MTS_EPI_SWIGLU: return launch_pair<MTS_EPI_SWIGLU>(ta, tb, p, tiles, stream);
    case MTS_EPI_ROPE_QK: return launch_pair<MTS_EPI_ROPE_QK>(ta, tb, p, tiles, stream);
    default: return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: unknown epilogue");
  }
}

}  // namespace mts
