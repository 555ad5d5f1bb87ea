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
//   warps 2..9      : epilogue      — tcgen05.ld 32x32b.x32 (thread = row, 32 columns), fused
//                     bias / residual-add / gelu_new / silu*up, vectorised global stores; two warps per
//                     TMEM lane quarter share the column chunks on the direct-store path (gemm_common.cuh)
//
// Algorithmic work per launch: 2*m*n*k*batch FLOP (roofline: tensor pipe).
#include <stdlib.h>
#include <string.h>

#include "gemm_common.cuh"

namespace mts {

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

// Work of one CTA: plain = whole tiles blockIdx.x, +gridDim.x, ... ; stream-K = the contiguous range of (tile, k-block)
// units [U*c/G, U*(c+1)/G) cut at tile boundaries.  All three warp roles walk the same sequence of segments.
struct SegIter {
  long long u, u_end;
  int k_blocks, step;
  bool streamk;
};
__device__ __forceinline__ SegIter seg_init(const GemmParams& p, int num_tiles, int k_blocks) {
  SegIter s;
  s.k_blocks = k_blocks;
  s.streamk = p.streamk != 0 || p.ksplit > 1;   // cluster split-K: grid = tiles x ksplit, the same even unit ranges
  if (s.streamk) {
    const long long U = (long long)num_tiles * k_blocks;
    s.u = U * blockIdx.x / gridDim.x;
    s.u_end = U * (blockIdx.x + 1) / gridDim.x;
    s.step = 0;
  } else {
    s.u = blockIdx.x;
    s.u_end = num_tiles;
    s.step = gridDim.x;
  }
  return s;
}
__device__ __forceinline__ bool seg_next(SegIter& s, int& tile, int& kb0, int& kb1) {
  if (s.u >= s.u_end) return false;
  if (s.streamk) {
    tile = (int)(s.u / s.k_blocks);
    kb0 = (int)(s.u - (long long)tile * s.k_blocks);
    const long long rem = s.u_end - s.u;
    kb1 = (long long)kb0 + rem < (long long)s.k_blocks ? (int)(kb0 + rem) : s.k_blocks;
    s.u += kb1 - kb0;
  } else {
    tile = (int)s.u;
    kb0 = 0;
    kb1 = s.k_blocks;
    s.u += s.step;
  }
  return true;
}

// TF32 = true: A and B are fp32 in global / shared memory and the tensor cores read them as TF32 (kind::tf32, K = 8 per
// instruction).  A 128-byte swizzled smem row then holds 32 elements instead of 64, so a stage covers half the K extent
// with the same bytes, the same descriptors and the same 4 MMAs; everything else is shared with the bf16 kernel.
template <int BN, int EPI, bool TF32 = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_nt_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmParams p,
                    const __grid_constant__ CUtensorMap tmap_a_lo,
                    const __grid_constant__ CUtensorMap tmap_b_lo) {
  using Cfg = GemmCfg<BN>;
  if (threadIdx.x == 0) GEMM_STAMP(0);
  constexpr int kStages = Cfg::kStages;
  constexpr int kBK = TF32 ? kBlockK / 2 : kBlockK;   // elements per 128-byte smem row
  // p.split3 (TF32 only): fp32-grade contraction from TF32 pieces, A = A_hi + A_lo, B = B_hi + B_lo (each piece
  // TF32-representable, mts_split_tf32): the k loop runs three times — A_hi B_hi, A_lo B_hi, A_hi B_lo — into the same
  // accumulator (the A_lo B_lo term is below fp32 resolution).  Only the producer's choice of tensor map changes.
  const int k_segments = (TF32 && p.split3) ? 3 : 1;

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
  // cluster split-K: peers_free = every peer's operand ring is idle (its MMAs are done), parts_in = every peer's partial of
  // this CTA's column part has landed in this CTA's ring
  const uint32_t peers_free_bar = bar_base + 8u * (2 * kStages + 5);
  const uint32_t parts_in_bar = bar_base + 8u * (2 * kStages + 6);
  const uint32_t stage_base = bar_base + Cfg::kBarrierBytes;  // 16-byte aligned
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (p.m + kBlockM - 1) / kBlockM;
  const int n_blocks = (p.n + BN - 1) / BN;
  const int k_blocks = (p.k + kBK - 1) / kBK;
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
    if (p.ksplit > 1) {
      mbar_init(peers_free_bar, p.ksplit - 1);                          // one arrive per peer
      mbar_init(parts_in_bar, (p.ksplit - 1) * (kNumEpiThreads / 32));   // one arrive per peer epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (k_segments > 1) {
      tma_prefetch_desc(&tmap_a_lo);
      tma_prefetch_desc(&tmap_b_lo);
    }
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  if (p.ksplit > 1) cluster_sync_all();   // the peers' barriers exist before anybody arrives on them remotely
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) GEMM_STAMP(1);
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; its results are
  // needed from here on.  The successor may be scheduled as soon as every CTA of this grid got this far.
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) GEMM_STAMP(2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it_ = seg_init(p, num_tiles, k_blocks);
      int tile, kb0, kb1;
      while (seg_next(it_, tile, kb0, kb1)) {
        const int b = tile / tiles_per_batch;
        const int t = tile - b * tiles_per_batch;
        int m_blk, n_blk;
        tile_coords(t, m_blocks, n_blocks, m_blk, n_blk);
        for (int seg = 0; seg < k_segments; ++seg) {
          const CUtensorMap* ma = (seg == 1) ? &tmap_a_lo : &tmap_a;
          const CUtensorMap* mb = (seg == 2) ? &tmap_b_lo : &tmap_b;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_3d(smem_a + stage * Cfg::kABytes, ma, full_bar(stage), kb * kBK,
                        m_blk * kBlockM, p.a_batched ? b : 0, kEvictNormal);
            tma_load_3d(smem_b + stage * Cfg::kBBytes, mb, full_bar(stage), kb * kBK,
                        n_blk * BN, p.b_batched ? b : 0, kEvictNormal);
            if (kb == kb0 && seg == 0) GEMM_STAMP(3);
            if (kb == kb1 - 1) GEMM_STAMP(9);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = TF32 ? umma_idesc_tf32(kBlockM, BN) : umma_idesc_bf16(kBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      SegIter it_ = seg_init(p, num_tiles, k_blocks);
      int tile, kb0, kb1;
      for (; seg_next(it_, tile, kb0, kb1); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 200 + acc);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        const int n_kb = (kb1 - kb0) * k_segments;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tc_fence_after();
          if (kb == 0) GEMM_STAMP(4);
          const uint64_t adesc = umma_desc_sw128(smem_a + stage * Cfg::kABytes);
          const uint64_t bdesc = umma_desc_sw128(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 (8 tf32) = 32 bytes inside the 128-byte swizzled row: +2 in (addr >> 4)
            if constexpr (TF32) umma_tf32(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else                umma_bf16(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        GEMM_STAMP(5);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9
    // TMEM -> registers (thread = accumulator row, 32 columns) -> per-warp padded smem tile ->
    // coalesced 16-byte global accesses (a quarter-warp covers one contiguous 128-byte row piece).
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int eset = (warp - 2) >> 2;                       // 0: warps 2..5 (own the staging buffers), 1: warps 6..9
    const int parts = epilogue_parts<EPI>(p);
    float* stage_buf = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw))) +
                       quarter * (32 * kEpiPitch);
    int it = 0;
    SegIter it_ = seg_init(p, num_tiles, k_blocks);
    int tile, kb0, kb1;
    for (; seg_next(it_, tile, kb0, kb1); ++it) {
      const int b = tile / tiles_per_batch;
      const int t = tile - b * tiles_per_batch;
      int m_blk, n_blk;
      tile_coords(t, m_blocks, n_blocks, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();
      if (threadIdx.x == 64) GEMM_STAMP(6);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
      if (p.ksplit > 1) {
        // ---- cluster split-K (one segment per CTA; EPI is STORE / RESID_ADD / GELU_NEW, batch 1)
        const int ks = p.ksplit, r = (int)cluster_ctarank();
        const int cols_per = BN / ks, cpp = cols_per / 32;            // columns / 32-column chunks of one part
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t part_bytes = (uint32_t)(kBlockM * cols_per * 4);
        // this CTA's MMAs are done (tfull): its operand ring is idle -> tell every peer; wait until all peers said so
        if (threadIdx.x == 64)
          for (int q = 0; q < ks; ++q)
            if (q != r) mbar_arrive_release_cluster(peers_free_bar, (uint32_t)q);
        mbar_wait_acquire_cluster(peers_free_bar, 0, 500);
        // send: chunk ci belongs to part ci / cpp; partial slot of sender r in owner q = r's index among q's peers.
        // Slot layout as the stream-K workspace: [chunk][group of 4 columns][row][4] fp32 (conflict-free 16-byte stores)
#pragma unroll 1
        for (int ci = eset; ci < BN / 32; ci += 2) {
          const int q = ci / cpp;
          if (q == r) continue;
          uint32_t rr[32];
          __syncwarp();
          tmem_ld_32x32(taddr + ci * 32, rr);
          tmem_ld_wait();
          const uint32_t local = smem_base + (uint32_t)(r < q ? r : r - 1) * part_bytes +
                                 (uint32_t)((((ci - q * cpp) * 8) * 128 + row_in_tile) * 16);
          const uint32_t remote = mapa_cluster(local, (uint32_t)q);
#pragma unroll
          for (int g = 0; g < 8; ++g) st_cluster_v4(remote + g * 2048, rr[4 * g], rr[4 * g + 1], rr[4 * g + 2], rr[4 * g + 3]);
        }
        fence_acq_rel_cluster();
        __syncwarp();
        if (lane == 0)
          for (int q = 0; q < ks; ++q)
            if (q != r) mbar_arrive_release_cluster(parts_in_bar, (uint32_t)q);
        mbar_wait_acquire_cluster(parts_in_bar, 0, 501);
        const int kparts = (parts == 2 && cpp >= 2) ? 2 : 1;
        if (eset < kparts) {
          const float* src = reinterpret_cast<const float*>(smem_raw + (smem_base - smem_u32(smem_raw))) + row_in_tile * 4;
          epilogue_tile<BN, EPI>(p, b, m_blk * kBlockM + quarter * 32, n_blk, taddr + r * cols_per, stage_buf, lane,
                                 r * cols_per, cols_per, eset, kparts, src, ks - 1, (int64_t)kBlockM * cols_per, true);
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        continue;
      }
      bool run_epilogue = eset < parts;
      const float* sk_src = nullptr;     // owner of a split tile: the contributors' partials, added inside the epilogue
      int sk_n = 0;
      const int64_t slot_elems = (int64_t)kBlockM * BN;
      if (p.streamk && !(kb0 == 0 && kb1 == k_blocks)) {
        // Workspace slot of a CTA: [32-column chunk][group of 4 columns][row][4] fp32 — for a fixed chunk and group the 128
        // rows (= lanes of the four warp quarters) are 16 bytes apart, so every warp access is one contiguous 512 bytes.
        const int row_in_tile = quarter * 32 + lane;
        if (kb0 > 0) {
          // ---- contributor: this CTA's range STARTS inside the tile (always its first segment under stream-K, its only one
          // under the even split, so the partial is parked without waiting for anybody).  Raw fp32 partial -> workspace
          // slot (both warp sets, alternate chunks), then the flag.
          float* ws = p.sk_ws + (int64_t)blockIdx.x * slot_elems + row_in_tile * 4;
#pragma unroll 1
          for (int ci = eset; ci < BN / 32; ci += 2) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld_32x32(taddr + ci * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<uint4*>(ws + (ci * 8 + j4) * 512) = make_uint4(r[4 * j4], r[4 * j4 + 1], r[4 * j4 + 2], r[4 * j4 + 3]);
          }
          __threadfence();
          asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight epilogue warps
          if (threadIdx.x == 64) st_release_gpu(p.sk_flags + blockIdx.x, p.sk_epoch);
          run_epilogue = false;
        } else if (eset < parts) {
          // ---- owner: holds the tile's FIRST k-blocks (the last segment of its range).  The CTAs that hold the rest — a
          // run of consecutive higher CTA indices — parked their partials; the epilogue adds them to the accumulator chunk
          // it has just read, in ascending CTA order (deterministic), before alpha / bias / activation.
          const long long U = (long long)num_tiles * k_blocks, G = gridDim.x;
          const long long last_unit = (long long)(tile + 1) * k_blocks - 1;
          int c_last = (int)(last_unit * G / U);
          while (U * (c_last + 1) / G <= last_unit) ++c_last;
          while (U * c_last / G > last_unit) --c_last;
          for (int cc = (int)blockIdx.x + 1; cc <= c_last; ++cc) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(p.sk_flags + cc) != p.sk_epoch) {
              if (clock64() - t0 > MTS_WATCHDOG_CYCLES) {
                printf("[mtsb200] watchdog: block %d stuck waiting for the stream-K partial of block %d\n", (int)blockIdx.x, cc);
                __trap();
              }
            }
          }
          sk_src = p.sk_ws + ((int64_t)blockIdx.x + 1) * slot_elems + row_in_tile * 4;
          sk_n = c_last - (int)blockIdx.x;
        }
      }
      if (run_epilogue)
        epilogue_tile<BN, EPI>(p, b, m_blk * kBlockM + quarter * 32, n_blk, taddr, stage_buf, lane, 0, BN, eset, parts,
                               sk_src, sk_n, slot_elems);
      if (sk_n > 0) {
        // every partial has exactly one reader: lower the flags again, so that the next launch — or the next replay of
        // a captured graph, which carries the same sk_epoch — starts from "not ready"
        if (parts == 2) asm volatile("bar.sync 2, 256;" ::: "memory");
        else            asm volatile("bar.sync 2, 128;" ::: "memory");
        if (threadIdx.x == 64)
          for (int cc = (int)blockIdx.x + 1; cc <= (int)blockIdx.x + sk_n; ++cc) st_release_gpu(p.sk_flags + cc, 0);
      }
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (threadIdx.x == 64) GEMM_STAMP(7);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) GEMM_STAMP(8);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// lo maps of the fp32-grade split (TF32 kernels with p.split3); the plain kernels get the main maps again (unused)
static thread_local const CUtensorMap* g_ta_lo = nullptr;
static thread_local const CUtensorMap* g_tb_lo = nullptr;

template <int BN, int EPI, bool TF32 = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       int num_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_nt_kernel<BN, EPI, TF32>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gemm smem)", e);
    attr_done = true;
  }
  // stream-K: one CTA per SM, each takes an equal share of the (tile, k-block) units
  const int grid = p.streamk ? p.sk_grid : (num_tiles < num_sms() ? num_tiles : num_sms());
  static int dbg_on = -1;
  static long long* dbg_buf = nullptr;
  if (dbg_on < 0) { const char* e = getenv("MTS_GEMM_DBG"); dbg_on = (e && e[0] == '1') ? 1 : 0; }
  GemmParams pp = p;
  pp.dbg = nullptr;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(long long), stream);
    pp.dbg = dbg_buf;
  }
  cudaError_t le;
  if (p.ksplit > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_tiles * p.ksplit);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.ksplit;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    le = cudaLaunchKernelEx(&cfg, kern, ta, tb, pp, (TF32 && g_ta_lo) ? *g_ta_lo : ta, (TF32 && g_tb_lo) ? *g_tb_lo : tb);
  } else {
    le = launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, pp,
                    (TF32 && g_ta_lo) ? *g_ta_lo : ta, (TF32 && g_tb_lo) ? *g_tb_lo : tb);
  }
  if (le != cudaSuccess) return set_cuda_error("cudaLaunchKernelEx(gemm_bf16_nt_kernel)", le);
  count_launch();
  if (dbg_on) {
    long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[gemm dbg] m=%d n=%d k=%d BN=%d epi=%d tiles=%d grid=%d: setup done=%lld pdl passed=%lld first TMA issued=%lld "
            "last TMA issued=%lld first stage landed=%lld MMAs issued=%lld accumulator seen=%lld epilogue done=%lld end=%lld\n",
            p.m, p.n, p.k, BN, EPI, num_tiles, grid, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[9] - h[0], h[4] - h[0], h[5] - h[0],
            h[6] - h[0], h[7] - h[0], h[8] - h[0]);
  }
  return check_launch("gemm_bf16_nt_kernel");
}

template <int BN, bool TF32 = false>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                        int num_tiles, cudaStream_t stream) {
  switch (epi) {
    case MTS_EPI_STORE: return launch_gemm<BN, MTS_EPI_STORE, TF32>(ta, tb, p, num_tiles, stream);
    case MTS_EPI_RESID_ADD: return launch_gemm<BN, MTS_EPI_RESID_ADD, TF32>(ta, tb, p, num_tiles, stream);
    case MTS_EPI_GELU_NEW: return launch_gemm<BN, MTS_EPI_GELU_NEW, TF32>(ta, tb, p, num_tiles, stream);
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

namespace mts {
void attn_tc_set(int v);
int launch_gemm_2cta(int epilogue, const mts_gemm_args* a, const GemmParams& p, cudaStream_t stream);
static int g_gemm_2cta = -1;
bool gemm_2cta_enabled() {
  if (g_gemm_2cta < 0) {
    const char* e = getenv("MTS_GEMM_2CTA");
    g_gemm_2cta = (e && e[0] == '0') ? 0 : 1;   // on by default: +5..18 % over the 1-CTA kernel on B200
  }
  return g_gemm_2cta == 1;
}
// 0 off (default) | 1 auto (uneven whole-tile schedules) | 2 whenever legal.  Off by default: measured on B200
// (profiles/r02_streamk_gemm.md) the contiguous unit ranges spread the concurrently live weight strips over the whole B
// matrix and the fix-up serialises at the end of every CTA; the plain schedule wins at every BASELINE shape today.
static int g_streamk = -1;
static int streamk_mode() {
  if (g_streamk < 0) {
    const char* e = getenv("MTS_STREAMK");
    g_streamk = e ? atoi(e) : 0;
    if (g_streamk < 0 || g_streamk > 3) g_streamk = 0;
  }
  return g_streamk;
}
static int g_epi_direct = -1;      // MTS_EPI_DIRECT=0 / mts_set_option("epi_direct", 0): always stage the epilogue through smem
static bool epi_direct_enabled() {
  if (g_epi_direct < 0) {
    const char* e = getenv("MTS_EPI_DIRECT");
    g_epi_direct = (e && e[0] == '0') ? 0 : 1;
  }
  return g_epi_direct == 1;
}
// cluster split-K of the single-CTA kernel: -1 auto (default) | 0 off | 2 / 4 forced whenever legal (experiments, tests)
static int g_gemm_ksplit = -2;
static int gemm_ksplit_mode() {
  if (g_gemm_ksplit == -2) {
    const char* e = getenv("MTS_GEMM_KSPLIT");
    g_gemm_ksplit = e ? atoi(e) : -1;
    if (!(g_gemm_ksplit == -1 || g_gemm_ksplit == 0 || g_gemm_ksplit == 2 || g_gemm_ksplit == 4)) g_gemm_ksplit = -1;
  }
  return g_gemm_ksplit;
}
static int g_gemm_force = 0;      // mts_set_option("gemm_force", 0 auto | 1 single-CTA kernel | 2 CTA-pair kernel): experiments
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("MTS_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;          // on by default
  }
  return g_pdl == 1;
}
}  // namespace mts

using namespace mts;

static_assert(sizeof(mts_gemm_args) == 248, "mts_gemm_args layout is part of the C ABI (ctypes mirror: _lib.GemmArgs)");

extern "C" int mts_set_option(const char* name, int value) {
  if (name && !strcmp(name, "gemm_2cta")) { g_gemm_2cta = value ? 1 : 0; return MTS_OK; }
  if (name && !strcmp(name, "pdl")) { g_pdl = value ? 1 : 0; return MTS_OK; }
  if (name && !strcmp(name, "gemm_force")) { g_gemm_force = value; return MTS_OK; }
  if (name && !strcmp(name, "gemm_ksplit")) {
    if (!(value == -1 || value == 0 || value == 2 || value == 4))
      return set_error(MTS_ERR_INVALID_ARG, "mts_set_option: gemm_ksplit is -1 (auto), 0, 2 or 4");
    g_gemm_ksplit = value;
    return MTS_OK;
  }
  if (name && !strcmp(name, "epi_direct")) { g_epi_direct = value ? 1 : 0; return MTS_OK; }
  if (name && !strcmp(name, "attn_tc")) { attn_tc_set(value); return MTS_OK; }
  if (name && !strcmp(name, "streamk")) { g_streamk = (value < 0 || value > 3) ? 0 : value; return MTS_OK; }
  return set_error(MTS_ERR_INVALID_ARG, "mts_set_option: unknown option '%s'", name ? name : "(null)");
}

extern "C" int mts_gemm(const mts_gemm_args* a, mts_stream_t stream_) {
  if (!a) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: null args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a->m <= 0 || a->n <= 0 || a->k <= 0 || a->batch <= 0)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: m, n, k, batch must be positive");
  if (!a->a || !a->b || !a->d) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: null operand");
  if (a->ab_dtype != MTS_BF16 && a->ab_dtype != MTS_F32)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: bad ab_dtype");
  const bool tf32 = a->ab_dtype == MTS_F32;
  const int ab_vec = tf32 ? 4 : 8;          // elements per 16 bytes of an operand row
  if ((a->lda % ab_vec) || (a->ldb % ab_vec) || a->lda < a->k || a->ldb < a->k)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: lda/ldb must be >= k and multiples of 16 bytes");
  if ((reinterpret_cast<uintptr_t>(a->a) & 15) || (reinterpret_cast<uintptr_t>(a->b) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: a and b must be 16-byte aligned");
  if ((a->a_batch_stride % ab_vec) || (a->b_batch_stride % ab_vec))
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: batch strides must be multiples of 16 bytes");
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
      if ((f32 && !tf32) || a->d_transposed)
        return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: GELU_NEW needs a non-transposed D (bf16; fp32 with fp32 operands)");
      break;
    case MTS_EPI_ROPE_QK:
      if ((f32 != tf32) || a->d_transposed || a->bias_axis != MTS_BIAS_NONE || !a->rope_cos || !a->rope_sin ||
          (a->rope_hd != 64 && a->rope_hd != 128) || a->rope_L <= 0 || a->rope_prefix < 0 || (a->rope_cols % a->rope_hd) ||
          (a->n % a->rope_hd) || (bn != 0 && bn != 256) || (n_store % 8) || (a->ldd % 8) || (a->d_batch_stride % 8) ||
          (reinterpret_cast<uintptr_t>(a->d) & 15) ||
          (reinterpret_cast<uintptr_t>(a->rope_cos) & 15) || (reinterpret_cast<uintptr_t>(a->rope_sin) & 15))
        return set_error(MTS_ERR_INVALID_ARG,
                         "mts_gemm: ROPE_QK needs D of the operand precision (bf16 / fp32), no bias, 16-byte aligned tables "
                         "and rows, head dim 64/128 dividing n and rope_cols, block_n 256");
      bn = 256;
      break;
    case MTS_EPI_SWIGLU:
      if ((f32 != tf32) || a->d_transposed || a->bias_axis != MTS_BIAS_NONE || (a->n % 256) ||
          (bn != 0 && bn != 256))
        return set_error(MTS_ERR_INVALID_ARG,
                         "mts_gemm: SWIGLU needs D of the operand precision (bf16 / fp32), no bias, n % 256 == 0 "
                         "(packed gate/up), block_n 256");
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
  // ---- stream-K: worth it when whole tiles leave the SMs unevenly loaded (small m) and every CTA still gets a few k-blocks
  int use_sk = 0, sk_grid = num_sms();
  {
    const int skm = streamk_mode();
    const int kb_elems = (a->ab_dtype == MTS_F32) ? kBlockK / 2 : kBlockK;
    if (skm != 0 && a->sk_workspace && a->sk_flags && a->batch == 1 && g_gemm_force != 2 &&
        (reinterpret_cast<uintptr_t>(a->sk_workspace) & 15) == 0 && a->sk_flags_len >= num_sms()) {
      int bn_sk = bn != 0 ? bn : (a->n >= 2048 ? 256 : 128);
      const long tiles = (long)((a->m + kBlockM - 1) / kBlockM) * ((a->n + bn_sk - 1) / bn_sk);
      const long kbs = (a->k + kb_elems - 1) / kb_elems;
      const long G = num_sms();
      const double eff = (double)tiles / (double)(((tiles + G - 1) / G) * G);
      const bool fits = a->sk_workspace_bytes >= (int64_t)G * kBlockM * bn_sk * 4;
      const bool enough = tiles * kbs >= 4 * G && kbs >= 4;
      if (skm == 3) {
        // even split-K: tiles x s CTAs in one wave, CTA c holds k range c % s of tile c / s (the general unit arithmetic with
        // G = tiles x s; the owner is the CTA with the first range, its s - 1 contributors are the next CTA indices)
        long sp = G / tiles;
        if (sp > 4) sp = 4;
        while (sp >= 2 && (kbs % sp) != 0) --sp;
        if (tiles >= 1 && sp >= 2 && kbs / sp >= 8 && fits) { use_sk = 1; bn = bn_sk; sk_grid = (int)(tiles * sp); }
      } else if (fits && enough && (skm == 2 || eff < 0.85)) { use_sk = 1; bn = bn_sk; }
    }
  }
  // ---- cluster split-K: few tiles with a deep k (GPT-2-sized projections on ~900 rows): clusters of s CTAs share a tile
  int ksplit = 0;
  {
    const int km = gemm_ksplit_mode();
    const bool epi_ok = a->epilogue == MTS_EPI_STORE || a->epilogue == MTS_EPI_RESID_ADD || a->epilogue == MTS_EPI_GELU_NEW;
    if (km != 0 && !use_sk && !tf32 && epi_ok && a->batch == 1 && !a->d_transposed && g_gemm_force != 2) {
      const long kbs = (a->k + kBlockK - 1) / kBlockK, mb = (a->m + kBlockM - 1) / kBlockM;
      auto legal = [&](int bn_k, int s) {
        const long tiles = mb * ((a->n + bn_k - 1) / bn_k);
        return tiles * s <= num_sms() && (kbs % s) == 0 && kbs / s >= 4 && bn_k / s >= 32;
      };
      if (km > 0) {                               // forced (tests / sweeps): the caller's tile width or the widest legal one
        const int bn_k = bn != 0 ? bn : (km == 4 ? 256 : 128);
        if ((bn_k == 64 || bn_k == 128 || bn_k == 256) && legal(bn_k, km)) { ksplit = km; bn = bn_k; }
      } else if (bn == 0) {
        // auto: measured on B200 (profiles/r02_ksplit_gemm.md): pairs on 128-wide tiles win whenever they fit one wave and
        // the k loop is deep enough to pay for the exchange (K >= 2048); the distributed-shared-memory stores run at
        // ~20 B/clk per SM, so clusters of four (three quarters of a 128 x 256 tile sent) stay behind pairs
        if (kbs >= 32 && legal(128, 2)) { ksplit = 2; bn = 128; }
      }
    }
  }
  if (bn == 0) bn = pick_block_n(a->m, a->n, a->batch);
  if (bn != 64 && bn != 128 && bn != 256)
    return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: block_n must be 0, 64, 128 or 256");

  const int a_batched = (a->batch > 1 && a->a_batch_stride != 0) ? 1 : 0;
  const int b_batched = (a->batch > 1 && a->b_batch_stride != 0) ? 1 : 0;
  CUtensorMap ta, tb;
  const int elem = tf32 ? 4 : 2, box_k = tf32 ? kBlockK / 2 : kBlockK;
  int rc = get_tmap_3d(&ta, a->a, a->k, a->m, a_batched ? a->batch : 1, a->lda,
                       a_batched ? a->a_batch_stride : (int64_t)a->m * a->lda, box_k, kBlockM, elem);
  if (rc) return rc;
  rc = get_tmap_3d(&tb, a->b, a->k, a->n, b_batched ? a->batch : 1, a->ldb,
                   b_batched ? a->b_batch_stride : (int64_t)a->n * a->ldb, box_k, bn, elem);
  if (rc) return rc;

  GemmParams p;
  p.dbg = nullptr;
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
  p.rope_cos = a->rope_cos; p.rope_sin = a->rope_sin;
  p.rope_L = a->rope_L; p.rope_hd = a->rope_hd; p.rope_cols = a->rope_cols; p.rope_prefix = a->rope_prefix;
  p.drop_thresh = 0; p.drop_scale = 1.0f; p.drop_seed = 0;
  if (a->drop_p != 0.0f) {
    if (a->epilogue != MTS_EPI_RESID_ADD || !(a->drop_p > 0.0f && a->drop_p < 1.0f))
      return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: drop_p needs the RESID_ADD epilogue and 0 <= p < 1");
    p.drop_thresh = (uint32_t)((double)a->drop_p * 4294967296.0);
    p.drop_scale = 1.0f / (1.0f - a->drop_p);
    p.drop_seed = a->drop_seed;
  }
  p.round_tf32 = (a->round_tf32 && f32) ? 1 : 0;
  {
    const int eb = f32 ? 4 : 2;
    p.direct = (epi_direct_enabled() && vec_ok && !a->d_transposed && (reinterpret_cast<uintptr_t>(a->d) & 31) == 0 &&
                ((a->ldd * eb) % 32) == 0 && ((a->d_batch_stride * eb) % 32) == 0 &&
                (!a->c || (reinterpret_cast<uintptr_t>(a->c) & 31) == 0))
                   ? 1 : 0;
  }
  p.precise = tf32 ? 1 : 0;
  p.streamk = use_sk;
  p.sk_grid = sk_grid;
  p.ksplit = ksplit;
  p.sk_ws = static_cast<float*>(a->sk_workspace);
  p.sk_flags = a->sk_flags;
  p.sk_epoch = a->sk_epoch;
  p.aux = nullptr; p.ld_aux = 0;
  if (a->aux && tf32) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: aux is a training (bf16) feature");
  if (a->aux) {
    if (a->epilogue != MTS_EPI_SWIGLU || a->batch != 1 || a->ld_aux < a->n || (a->ld_aux % 8) ||
        (reinterpret_cast<uintptr_t>(a->aux) & 15))
      return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: aux needs the SWIGLU epilogue, batch 1, ld_aux >= n, 16-byte rows");
    p.aux = static_cast<__nv_bfloat16*>(a->aux); p.ld_aux = a->ld_aux;
  }

  p.split3 = 0;
  CUtensorMap ta_lo, tb_lo;
  g_ta_lo = g_tb_lo = nullptr;
  if (a->a_lo || a->b_lo) {
    if (!tf32 || !a->a_lo || !a->b_lo || (reinterpret_cast<uintptr_t>(a->a_lo) & 15) || (reinterpret_cast<uintptr_t>(a->b_lo) & 15))
      return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: a_lo / b_lo need fp32 operands, both given, 16-byte aligned");
    rc = get_tmap_3d(&ta_lo, a->a_lo, a->k, a->m, a_batched ? a->batch : 1, a->lda,
                     a_batched ? a->a_batch_stride : (int64_t)a->m * a->lda, box_k, kBlockM, elem);
    if (rc) return rc;
    rc = get_tmap_3d(&tb_lo, a->b_lo, a->k, a->n, b_batched ? a->batch : 1, a->ldb,
                     b_batched ? a->b_batch_stride : (int64_t)a->n * a->ldb, box_k, bn, elem);
    if (rc) return rc;
    g_ta_lo = &ta_lo; g_tb_lo = &tb_lo;
    p.split3 = 1;
  }
  if (tf32) {
    // evaluation parity mode (fp32 operands as TF32): single-CTA kernel only
    const long tl = (long)((a->m + kBlockM - 1) / kBlockM) * ((a->n + bn - 1) / bn) * a->batch;
    if (tl > 0x7fffffffL) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: too many tiles");
    if (a->epilogue == MTS_EPI_SWIGLU) return launch_gemm<256, MTS_EPI_SWIGLU, true>(ta, tb, p, (int)tl, stream);
    if (a->epilogue == MTS_EPI_ROPE_QK) return launch_gemm<256, MTS_EPI_ROPE_QK, true>(ta, tb, p, (int)tl, stream);
    switch (bn) {
      case 256: return dispatch_epi<256, true>(a->epilogue, ta, tb, p, (int)tl, stream);
      case 128: return dispatch_epi<128, true>(a->epilogue, ta, tb, p, (int)tl, stream);
      default: return dispatch_epi<64, true>(a->epilogue, ta, tb, p, (int)tl, stream);
    }
  }
  if (bn == 256 && !a->d_transposed && gemm_2cta_enabled() && !use_sk && ksplit == 0) {
    // CTA pairs own 256x256 tiles (74 pairs); a tile costs ~0.92 of two 128x256 tiles' time.  Prefer them
    // unless the 256-row granularity wastes more than it saves (odd / single 128-row block counts).
    const long nb256 = (a->n + 255) / 256;
    const long t1 = (long)((a->m + 127) / 128) * nb256 * a->batch, tp = (long)((a->m + 255) / 256) * nb256 * a->batch;
    const long cost1 = ((t1 + num_sms() - 1) / num_sms()) * 100;
    const long costp = ((tp + num_sms() / 2 - 1) / (num_sms() / 2)) * 92;
    if (g_gemm_force == 2 || (g_gemm_force == 0 && costp <= cost1)) return launch_gemm_2cta(a->epilogue, a, p, stream);
  }

  const int mb = (a->m + kBlockM - 1) / kBlockM;
  const int nb = (a->n + bn - 1) / bn;
  const long tiles_l = (long)mb * nb * a->batch;
  if (tiles_l > 0x7fffffffL) return set_error(MTS_ERR_INVALID_ARG, "mts_gemm: too many tiles");
  const int tiles = (int)tiles_l;

  if (a->epilogue == MTS_EPI_SWIGLU) return launch_gemm<256, MTS_EPI_SWIGLU>(ta, tb, p, tiles, stream);
  if (a->epilogue == MTS_EPI_ROPE_QK) return launch_gemm<256, MTS_EPI_ROPE_QK>(ta, tb, p, tiles, stream);
  switch (bn) {
    case 256: return dispatch_epi<256>(a->epilogue, ta, tb, p, tiles, stream);
    case 128: return dispatch_epi<128>(a->epilogue, ta, tb, p, tiles, stream);
    default: return dispatch_epi<64>(a->epilogue, ta, tb, p, tiles, stream);
  }
}
