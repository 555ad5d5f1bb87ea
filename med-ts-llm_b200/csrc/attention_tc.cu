// attention_tc.cu — causal attention forward for the short sequences of this path (prompt + patches, L <= 256) on the
// 5th-generation tensor cores.  Reference op: HF eager attention, HF:models/llama/modeling_llama.py:199-221 and
// HF:models/gpt2/modeling_gpt2.py:54-72 (softmax(Q K^T / sqrt(hd) + causal mask) V), q / k already rotated.
//
// One job = up to 128 query rows that are contiguous in the row layout (several whole samples' own tokens, or one
// 128-row tile of a longer sequence) against at most 256 keys, so the whole score tile lives in tensor memory and the
// softmax is a single pass (no running maximum, no rescaling of the output):
//
//   TMA      Q [128 x hd], K and V [keys x hd] (16-row boxes, 128B swizzle) -> shared memory
//   tcgen05  S[128 x keys] = Q K^T            (kind::f16, fp32 accumulators: 256 TMEM columns)
//   4 warps  thread = query row: row maximum and exp2 straight out of TMEM (tcgen05.ld), P (bf16) written into
//            128B-swizzled K-major shared memory, row sum kept in a register
//   tcgen05  O[128 x hd] = P V                (V consumed MN-major, exactly as TMA delivers its rows; 128 TMEM columns)
//   4 warps  O / rowsum -> bf16 rows in global memory (+ log-sum-exp for the backward)
//
// Key columns of a job: [keys every row sees: the shared prompt prefix] followed by one 16-aligned column segment per
// sample for (the earlier tiles of the same sequence and) its own, causally masked keys; a segment starts with
// Lc mod 16 dummy columns (its first 16-row box simply begins that many rows early).  Every key therefore sits at a column
// that is congruent to its POSITION modulo 16 and in position order, masked columns contribute exact zeros and the row
// sum runs strictly left to right: a row's result does not depend on which other samples share its tile, nor on whether
// the prompt prefix is stored once (shared-prefix layout) or per sample — bit-identical outputs either way.
//
// A CTA (one per SM, persistent) walks a contiguous range of jobs, head-major, so that the prefix K / V of a head stay
// resident in shared memory across its jobs.  Warp roles: 0-3 softmax, 4-7 epilogue (both sets cover the four TMEM lane
// quarters), 8 TMA producer, 9 MMA issuer.  The S MMA of job i+1 is issued right behind the O MMA of job i, so softmax
// i+1 runs while the epilogue warps drain O of job i: the kernel's bound is the TMEM read bandwidth (S twice, O once).
//
// Algorithmic work per launch: 4 * hd * (visible query-key pairs) FLOP (roofline: tensor pipe).
#include <stdlib.h>

#include "mts_internal.h"
#include "ptx.cuh"

namespace mts {

constexpr int kTcThreads = 320;   // warps 0-3 softmax, 4-7 epilogue, 8 TMA producer, 9 MMA issuer

struct AttnTcParams {
  __nv_bfloat16* out;
  float* lse;
  int Bp, Lc, Ls, H, D;          // Lc = 0: plain layout (Ls = L positions per sample)
  int n_pt;                      // 128-row tiles of the shared prefix (per head)
  int n_qt;                      // 128-row tiles per sample; 1: several samples share a job
  int spj;                       // samples per job (n_qt == 1)
  int jobs_per_head, n_jobs;
  float scale_log2e;
  long long* dbg;                // MTS_ATTN_TC_DBG=1: clock64 stamps of block 0 (debug)
};
#define TC_STAMP(slot) do { if (p.dbg && blockIdx.x == 0) p.dbg[slot] = clock64(); } while (0)

struct TcJob {
  int head, q_row0, R;           // query rows: global rows [q_row0, q_row0 + R)
  int Ls, nseg;                  // rows per sample, samples in the job
  int na, na16;                  // prefix keys every row sees: global rows [0, na) in columns [0, na); na16 = na rounded up to 16
  int nb;                        // earlier own rows of the (single) sample, also seen by every row: global rows [q_row0 - nb, q_row0)
  int e, Lsp;                    // segment of sample s: columns [na16 + s*Lsp, +Lsp) <- global rows q_row0 - nb + s*Ls - e + [0, Lsp);
                                 // its keys start e columns in (e = Lc mod 16 keeps column = position modulo 16)
  int nk;                        // key columns in all (multiple of 16, <= 256)
  int64_t lse_off;               // lse index of local row r: lse_off + (r / Ls) * lse_stride + r % Ls
  int lse_stride;
};

// Jobs are numbered head-major; a CTA walks a contiguous range, so (head, index within the head) advance incrementally
// (one division per CTA, none per job).
struct TcJobIter {
  int head, i;
  __device__ __forceinline__ void init(const AttnTcParams& p, int job) { head = job / p.jobs_per_head; i = job - head * p.jobs_per_head; }
  __device__ __forceinline__ void next(const AttnTcParams& p) { if (++i == p.jobs_per_head) { i = 0; ++head; } }
};

__device__ __forceinline__ TcJob tc_decode(const AttnTcParams& p, const TcJobIter& it) {
  TcJob j;
  j.head = it.head;
  int i = it.i;
  j.nb = 0; j.lse_stride = 0; j.e = 0;
  if (i < p.n_pt) {                        // tile i of the prefix: an ordinary causal sequence of Lc positions
    j.q_row0 = 128 * i;
    j.R = min(128, p.Lc - 128 * i);
    j.Ls = j.R; j.nseg = 1;
    j.na = 0; j.nb = 128 * i;
    j.lse_off = (int64_t)j.head * p.Lc + 128 * i;
  } else {
    i -= p.n_pt;
    const int64_t lse_s = (int64_t)p.H * p.Lc;          // the samples' part follows the prefix part
    j.na = p.Lc;
    j.e = p.Lc & 15;
    if (p.n_qt == 1) {                     // samples b0s .. b0s + nseg - 1, all their own rows
      const int b0s = i * p.spj;
      j.nseg = min(p.spj, p.Bp - b0s);
      j.q_row0 = p.Lc + b0s * p.Ls;
      j.R = j.nseg * p.Ls;
      j.Ls = p.Ls;
      j.lse_off = lse_s + ((int64_t)b0s * p.H + j.head) * p.Ls;
      j.lse_stride = p.H * p.Ls;
    } else {                               // tile t of sample b: its earlier own rows are keys every row of the tile sees
      const int b = i / p.n_qt, t = i - b * p.n_qt;
      j.nseg = 1;
      j.q_row0 = p.Lc + b * p.Ls + 128 * t;
      j.R = min(128, p.Ls - 128 * t);
      j.Ls = j.R;
      j.nb = 128 * t;
      j.lse_off = lse_s + ((int64_t)b * p.H + j.head) * p.Ls + 128 * t;
    }
  }
  j.na16 = (j.na + 15) & ~15;
  j.Lsp = j.nb + ((j.e + j.Ls + 15) & ~15);
  j.nk = j.na16 + j.nseg * j.Lsp;
  return j;
}

// 2^x on the special-function unit, one instruction (results below 2^-126 flush to zero: far below bf16 resolution of P)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Bit j set: column c0 + j is visible to a row that sees the prefix columns [0, n_prefix) and its segment keys
// [lo, hi] (hi < lo: none).  One mask per 32-column chunk instead of three comparisons per element.
__device__ __forceinline__ uint32_t chunk_mask(int c0, int n_prefix, int lo, int hi) {
  const int np = min(max(n_prefix - c0, 0), 32);
  uint32_t m = np >= 32 ? 0xffffffffu : ((1u << np) - 1u);
  const int a = max(lo - c0, 0), b = min(hi - c0, 31);
  if (b >= a) m |= (0xffffffffu >> (31 - b)) & (0xffffffffu << a);
  return m;
}

// K / V rows [grow, grow + n) (n a multiple of 16) of one 64-column block -> shared-memory rows [srow, srow + n) of a slab, in
// as few TMA operations as the three box heights allow (a TMA instruction costs the issuing thread ~120 cycles whatever
// its size: 16-row boxes alone made the producer the slowest role of the backward).
__device__ __forceinline__ void tma_load_rows(uint32_t slab, const CUtensorMap* m128, const CUtensorMap* m64,
                                              const CUtensorMap* m16, uint32_t bar, int col, int srow, int grow, int n,
                                              uint64_t policy) {
  int r = 0;
  for (; r + 128 <= n; r += 128) tma_load_3d(slab + (srow + r) * 128, m128, bar, col, grow + r, 0, policy);
  for (; r + 64 <= n; r += 64) tma_load_3d(slab + (srow + r) * 128, m64, bar, col, grow + r, 0, policy);
  for (; r < n; r += 16) tma_load_3d(slab + (srow + r) * 128, m16, bar, col, grow + r, 0, policy);
}

// The same boxes as an L2 prefetch (no shared-memory destination, no barrier): issued for the NEXT job while the current one
// computes, so that its loads — which can only start when the current job's MMAs release the tiles — hit L2 instead of HBM.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_rows(const CUtensorMap* m128, const CUtensorMap* m64, const CUtensorMap* m16,
                                                  int col, int grow, int n) {
  int r = 0;
  for (; r + 128 <= n; r += 128) tma_prefetch_3d(m128, col, grow + r, 0);
  for (; r + 64 <= n; r += 64) tma_prefetch_3d(m64, col, grow + r, 0);
  for (; r < n; r += 16) tma_prefetch_3d(m16, col, grow + r, 0);
}

// 32-byte global store (STG.256): one full sector per lane
__device__ __forceinline__ void st_global_v8(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                             uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4),
               "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

// MN-major, 128B-swizzled shared-memory matrix descriptor: rows of the K dimension (keys) are 128 bytes (64 bf16 of the
// MN dimension) apart, 8-row groups 1024 bytes (SBO), the next 64 elements of the MN dimension `lbo` bytes further.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int HD>
struct AttnTcCfg {
  static constexpr int KB = HD / 64;                  // 64-column blocks of the head dim
  static constexpr int kSlabQ = 128 * 128;            // bytes: 128 query rows x 128 B
  static constexpr int kSlabKV = 256 * 128;           // 256 key rows x 128 B
  static constexpr int kSlabP = 128 * 128;            // 128 query rows x 64 keys
  static constexpr int kOffK = KB * kSlabQ;
  static constexpr int kOffV = kOffK + KB * kSlabKV;
  static constexpr int kOffP = kOffV + KB * kSlabKV;
  static constexpr int kOffLM = kOffP + 4 * kSlabP;   // row sums and scaled row maxima, softmax -> epilogue warps: 2 x 128 floats
  static constexpr int kOffBar = kOffLM + 1024;
  static constexpr int kSmemBytes = kOffBar + 128 + 1024;
};

template <int HD>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const __grid_constant__ CUtensorMap tmap_kv64, const AttnTcParams p) {
  using Cfg = AttnTcCfg<HD>;
  constexpr int KB = Cfg::KB;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = tc_smem_raw + (base - smem_u32(tc_smem_raw));
  const uint32_t sQ = base, sK = base + Cfg::kOffK, sV = base + Cfg::kOffV, sP = base + Cfg::kOffP;
  const uint32_t bar = base + Cfg::kOffBar;
  const uint32_t q_full = bar, v_full = bar + 8, q_empty = bar + 16, v_empty = bar + 24, s_full = bar + 32,
                 p_full = bar + 40, o_full = bar + 48, tmem_slot = bar + 56, o_empty = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(v_full, 1); mbar_init(q_empty, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_empty, 128);
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_kv); tma_prefetch_desc(&tmap_kv64); }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(base_ptr + Cfg::kOffBar + 56);
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) TC_STAMP(0);

  const int j_begin = (int)((long long)p.n_jobs * blockIdx.x / gridDim.x);
  const int j_end = (int)((long long)p.n_jobs * (blockIdx.x + 1) / gridDim.x);

  if (warp == 8) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ph = 0;
      int res_head = -1, res_na = -1;            // prefix rows [0, res_na) of head res_head are resident in the slabs
      TcJobIter it;
      it.init(p, j_begin);
      for (int job = j_begin; job < j_end; ++job, ph ^= 1u, it.next(p)) {
        const TcJob jb = tc_decode(p, it);
        const bool reuse = (jb.head == res_head && jb.na == res_na);
        const int col_h = jb.head * HD;
        const int n_boxes = (reuse ? 0 : (jb.na16 >> 4)) + jb.nseg * (jb.Lsp >> 4);
        const int seg_row0 = jb.q_row0 - jb.nb - jb.e;            // >= 0: sample jobs start at or after row Lc >= e
        auto load_keys = [&](uint32_t slab, int col0, uint32_t full) {
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t dst = slab + kb * Cfg::kSlabKV;
            const int col = col0 + col_h + kb * 64;
            if (!reuse) tma_load_rows(dst, &tmap_q, &tmap_kv64, &tmap_kv, full, col, 0, 0, jb.na16, kEvictLast);
            for (int s = 0; s < jb.nseg; ++s)
              tma_load_rows(dst, &tmap_q, &tmap_kv64, &tmap_kv, full, col, jb.na16 + s * jb.Lsp, seg_row0 + s * jb.Ls, jb.Lsp,
                            kEvictNormal);
          }
        };
        // Q and K: free once the S MMA of the previous job has retired
        mbar_wait(q_empty, ph ^ 1u, 500);
        mbar_arrive_expect_tx(q_full, KB * (Cfg::kSlabQ + n_boxes * 2048));
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d(sQ + kb * Cfg::kSlabQ, &tmap_q, q_full, col_h + kb * 64, jb.q_row0, 0, kEvictFirst);
        load_keys(sK, p.D, q_full);
        if (job == j_begin) TC_STAMP(1);
        // V: free once the O MMA of the previous job has retired
        mbar_wait(v_empty, ph ^ 1u, 501);
        mbar_arrive_expect_tx(v_full, KB * n_boxes * 2048);
        load_keys(sV, 2 * p.D, v_full);
        if (job == j_begin) TC_STAMP(2);
        res_head = jb.head;
        res_na = jb.na;
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ph = 0;
      const uint32_t tS = tmem_base, tO = tmem_base + 256;
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD) | (1u << 16);      // B (= V) is MN-major
      // S of job i+1 is issued BEFORE O of job i (both become possible when the softmax warps hand over P of job i and
      // release S): the softmax warps get their next score tile one MMA earlier and O of job i runs underneath pass 1.
      auto issue_s = [&](int nk, uint32_t ph_q) {
        mbar_wait(q_full, ph_q, 510);
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_bf16(128, (uint32_t)nk);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t adesc = umma_desc_sw128(sQ + kb * Cfg::kSlabQ);
          const uint64_t bdesc = umma_desc_sw128(sK + kb * Cfg::kSlabKV);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS, adesc + 2u * k, bdesc + 2u * k, idesc_s, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(q_empty);
        umma_commit(s_full);
      };
      TcJobIter it;
      it.init(p, j_begin);
      TcJob jb = tc_decode(p, it);
      if (j_begin < j_end) issue_s(jb.nk, 0);
      TC_STAMP(4);
      for (int job = j_begin; job < j_end; ++job, ph ^= 1u) {
        const int nk = jb.nk;                                                  // key columns, multiple of 16, <= 256
        it.next(p);
        TcJob jn = jb;
        if (job + 1 < j_end) jn = tc_decode(p, it);      // (before the wait: off the critical path)
        asm volatile("" ::"r"(jn.nk));
        mbar_wait(p_full, ph, 511);          // P is in shared memory and the softmax warps are done with S
        if (job + 1 < j_end) issue_s(jn.nk, ph ^ 1u);
        mbar_wait(v_full, ph, 512);
        mbar_wait(o_empty, ph ^ 1u, 513);    // the epilogue warps have drained O of the previous job
        tc_fence_after();
        if (job == j_begin) TC_STAMP(5);
        for (int ks = 0; ks < (nk >> 4); ++ks) {
          const uint64_t adesc = umma_desc_sw128(sP + (ks >> 2) * Cfg::kSlabP) + 2u * (ks & 3);
          const uint64_t bdesc = umma_desc_mn_sw128(sV + ks * 2048, Cfg::kSlabKV);
          umma_bf16(tO, adesc, bdesc, idesc_o, ks != 0 ? 1u : 0u);
        }
        umma_commit(v_empty);
        umma_commit(o_full);
        if (job == j_begin) TC_STAMP(6);
        jb = jn;
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------------------------------ softmax
    const int r = warp * 32 + lane;                               // query row of the tile = TMEM lane
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* p_row = base_ptr + Cfg::kOffP + r * 128;
    float* lm = reinterpret_cast<float*>(base_ptr + Cfg::kOffLM);
    const int rx = r & 7;
    uint32_t ph = 0;
    const int r_div = r / p.Ls;                    // sample of this row inside a several-samples job (the only job kind with s > 0)
    TcJobIter it;
    it.init(p, j_begin);
    for (int job = j_begin; job < j_end; ++job, ph ^= 1u, it.next(p)) {
      const TcJob jb = tc_decode(p, it);
      const int nk = jb.nk;
      const int nab = jb.na;                       // columns [0, na): the prefix, seen by every row
      const bool row_valid = r < jb.R;
      const int s = (row_valid && jb.nseg > 1) ? r_div : 0, t = r - s * jb.Ls;
      // this row's segment keys (earlier tiles of its sequence + own keys up to itself): columns [own_lo, own_hi];
      // rows past the tile see nothing
      const int own_lo = row_valid ? jb.na16 + s * jb.Lsp + jb.e : (1 << 30);
      const int own_hi = row_valid ? own_lo + jb.nb + t : -1;
      const int nab_row = row_valid ? nab : 0;
      // (keep the job arithmetic above the wait: the compiler would otherwise sink it behind the barrier, onto the critical path)
      asm volatile("" ::"r"(own_lo), "r"(own_hi), "r"(nab_row), "r"(nk), "r"(t));
      // S of this job is complete
      if (lane == 0) mbar_wait(s_full, ph, 520);
      __syncwarp();
      tc_fence_after();
      if (job == j_begin && threadIdx.x == 0) TC_STAMP(7);
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < nk; c0 += 32) {
        if (c0 >= nab && !__any_sync(0xffffffffu, c0 <= own_hi && c0 + 31 >= own_lo)) continue;   // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tS + c0, v);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // four chains: the maximum is order-independent
        if (c0 + 32 <= nab) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
        } else {
          const uint32_t bits = chunk_mask(c0, nab_row, own_lo, own_hi);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            m4[j & 3] = fmaxf(m4[j & 3], (bits & (1u << j)) ? __uint_as_float(v[j]) : -INFINITY);
        }
        mx = fmaxf(fmaxf(mx, fmaxf(m4[0], m4[1])), fmaxf(m4[2], m4[3]));
      }
      const float m_sc = row_valid ? mx * p.scale_log2e : 0.0f;
      if (job == j_begin && threadIdx.x == 0) TC_STAMP(8);
      // P is rewritten from here on: O of the previous job (issued after this job's S) must have finished reading it
      if (lane == 0) mbar_wait(o_full, ph ^ 1u, 523);
      __syncwarp();
      float l = 0.0f;
#pragma unroll 1
      for (int c0 = 0; c0 < nk; c0 += 32) {
        uint8_t* dst = p_row + (c0 >> 6) * Cfg::kSlabP;
        const int q0 = (c0 & 63) >> 3;
        if (c0 >= nab && !__any_sync(0xffffffffu, c0 <= own_hi && c0 + 31 >= own_lo)) {
#pragma unroll
          for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(dst + (((q0 + g) ^ rx) << 4)) = make_uint4(0, 0, 0, 0);
          continue;
        }
        uint32_t v[32];
        tmem_ld_32x32(tS + c0, v);
        tmem_ld_wait();
        float pv[32];
        if (c0 + 32 <= nab) {
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -m_sc));
        } else {
          // branch-free: the exponential is taken of every column (a masked one may overflow to +inf, never NaN) and
          // the mask selects afterwards — a per-element branch here diverges 32 times per chunk
          const uint32_t bits = chunk_mask(c0, nab_row, own_lo, own_hi);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float ev = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -m_sc));
            pv[j] = (bits & (1u << j)) ? ev : 0.0f;
          }
        }
        // the row sum runs strictly left to right (masked columns add exact zeros): independent of the tile's make-up
#pragma unroll
        for (int j = 0; j < 32; ++j) l += pv[j];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(dst + (((q0 + g) ^ rx) << 4)) =
              make_uint4(pack_bf16(pv[8 * g], pv[8 * g + 1]), pack_bf16(pv[8 * g + 2], pv[8 * g + 3]),
                         pack_bf16(pv[8 * g + 4], pv[8 * g + 5]), pack_bf16(pv[8 * g + 6], pv[8 * g + 7]));
      }
      // row sum and maximum go to the epilogue warps through shared memory, once they have read the previous job's
      if (lane == 0) mbar_wait(o_empty, ph ^ 1u, 522);
      __syncwarp();
      lm[r] = l;
      lm[128 + r] = m_sc;
      fence_proxy_async_smem();             // generic-proxy writes of P -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
      if (job == j_begin && threadIdx.x == 0) TC_STAMP(9);
      if (job == j_begin && lane == 0) TC_STAMP(16 + warp);
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue (warps 4..7)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t tO = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256;
    const float* lm = reinterpret_cast<const float*>(base_ptr + Cfg::kOffLM);
    uint32_t ph = 0;
    const int r_div = r / p.Ls;
    TcJobIter it;
    it.init(p, j_begin);
    for (int job = j_begin; job < j_end; ++job, ph ^= 1u, it.next(p)) {
      const TcJob jb = tc_decode(p, it);
      const bool row_valid = r < jb.R;
      const int s = (row_valid && jb.nseg > 1) ? r_div : 0, t = r - s * jb.Ls;
      __nv_bfloat16* orow = p.out + (int64_t)(jb.q_row0 + r) * p.D + (int64_t)jb.head * HD;
      float* lse_ptr = (p.lse != nullptr && row_valid) ? p.lse + jb.lse_off + (int64_t)s * jb.lse_stride + t : nullptr;
      asm volatile("" ::"l"(orow), "l"(lse_ptr));
      if (lane == 0) mbar_wait(o_full, ph, 521);
      __syncwarp();
      tc_fence_after();
      if (job == j_begin && threadIdx.x == 128) TC_STAMP(10);
      const float l = lm[r], m_sc = lm[128 + r];
      const float inv = l > 0.0f ? 1.0f / l : 0.0f;
      uint32_t v[HD / 32][32];
#pragma unroll
      for (int ci = 0; ci < HD / 32; ++ci) tmem_ld_32x32(tO + ci * 32, v[ci]);
      tmem_ld_wait();
      tc_fence_before();                    // O and (l, m) are in registers: hand both back before the global stores
      mbar_arrive(o_empty);
      if (row_valid) {
#pragma unroll
        for (int ci = 0; ci < HD / 32; ++ci) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {         // 16 bf16 = one full 32-byte sector per store
            uint32_t w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              w[q] = pack_bf16(__uint_as_float(v[ci][16 * g + 2 * q]) * inv, __uint_as_float(v[ci][16 * g + 2 * q + 1]) * inv);
            st_global_v8(orow + ci * 32 + 16 * g, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
          }
        }
      }
      if (lse_ptr != nullptr) *lse_ptr = (m_sc + log2f(l)) * 0.6931471805599453f;
      if (threadIdx.x == 128) TC_STAMP(job == j_begin ? 11 : 12);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(13);
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =============================================================================================================
// Backward on the shared-prefix layout with a frozen backbone (gradient enters at the samples' own rows and only their
// own q / k / v rows receive one): the same job structure as the forward, five tcgen05 contractions per job and no
// data movement between them — every operand is consumed in the layout TMA (or the previous stage) left it in:
//
//   S  = Q K^T,  dP = dO V^T            K-major A and B                          -> 2 x 256 TMEM columns
//   8 warps (two per TMEM lane quarter, alternating 32-column chunks; thread = query row):
//        P = exp2(S c - lse),  dS = P (dP - delta) scale,  delta = rowsum(dO o O)
//        dS (bf16) -> 128B-swizzled shared memory over the dead V tile, P of the own columns -> its own tile
//   dQ = dS K                           A = dS K-major, B = K MN-major            -> TMEM columns   0..127
//   dV = P_own^T dO                     A = P  MN-major (M = own keys), B = dO MN-major ->        128..255
//   dK = dS_own^T Q                     A = dS MN-major,                B = Q  MN-major ->        256..383
//   8 warps: TMEM -> (RoPE rotated back on dQ / dK) -> bf16 rows of dqkv
//
// Own-key column segments start at a multiple of 64 (the M block of dK / dV must begin on a swizzle row), 16-aligned per
// sample; no dummy columns (the backward makes no bit-identity promise between row layouts).
// Algorithmic work per launch: 10 * hd * (visible query-key pairs) FLOP.
// =============================================================================================================
struct AttnTcBwdParams {
  const __nv_bfloat16* out_own;      // [Bp*Ls, D]
  const float* lse_own;              // [Bp, H, Ls]
  const float* rope_cos;             // [>= Lc + Ls, hd/2] or null
  const float* rope_sin;
  __nv_bfloat16* dqkv_own;           // [Bp*Ls, 3D]
  float* delta_own;                  // [Bp, H, Ls] or null: rowsum(dO o O), kept for the prefix-key kernels of the full backward
  int Bp, Lc, Ls, H, D;
  int spj, groups, n_jobs;
  int na64, Lsp;                     // first own-key column (Lc rounded up to 64), column segment per sample
  float scale, scale_log2e;
  long long* dbg;                    // MTS_ATTN_TC_DBG=1: clock64 stamps of block 0, second job (debug)
};

template <int HD>
struct AttnTcBwdCfg {
  static constexpr int KB = HD / 64;
  static constexpr int kSlabQ = 128 * 128;
  static constexpr int kSlabKV = 256 * 128;
  static constexpr int kSlab = 128 * 128;                 // one 64-column slab of P / dS
  static constexpr int kOffK = KB * kSlabQ;
  static constexpr int kOffV = kOffK + KB * kSlabKV;
  static constexpr int kOffdO = kOffV + KB * kSlabKV;
  static constexpr int kOffdS = (HD == 128) ? kOffV : kOffdO + KB * kSlabQ;       // hd 128: over the V tile, dead after dP
  static constexpr int kOffP = (HD == 128) ? kOffdO + KB * kSlabQ : kOffdS + 4 * kSlab;
  static constexpr int kOffBar = kOffP + 2 * kSlab;
  static constexpr int kSmemBytes = kOffBar + 128 + 1024;
};

constexpr int kTcBwdThreads = 320;     // warps 0-7 compute (quarter = warp % 4, half = warp / 4), 8 TMA producer, 9 MMA issuer

template <int HD>
__global__ void __launch_bounds__(kTcBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const __grid_constant__ CUtensorMap tmap_kv64, const __grid_constant__ CUtensorMap tmap_do,
                   const AttnTcBwdParams p) {
  using Cfg = AttnTcBwdCfg<HD>;
  constexpr int KB = Cfg::KB;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = tc_smem_raw + (base - smem_u32(tc_smem_raw));
  const uint32_t sQ = base, sK = base + Cfg::kOffK, sV = base + Cfg::kOffV, sdO = base + Cfg::kOffdO,
                 sdS = base + Cfg::kOffdS, sP = base + Cfg::kOffP;
  const uint32_t bar = base + Cfg::kOffBar;
  const uint32_t ld_full = bar, smem_free = bar + 8, s_full = bar + 16, ds_full = bar + 24, g_full = bar + 32,
                 epi_done = bar + 40, tmem_slot = bar + 48;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(ld_full, 1); mbar_init(smem_free, 1); mbar_init(s_full, 1); mbar_init(ds_full, 256);
    mbar_init(g_full, 1); mbar_init(epi_done, 256);
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_kv); tma_prefetch_desc(&tmap_kv64); tma_prefetch_desc(&tmap_do);
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(base_ptr + Cfg::kOffBar + 48);
  pdl_wait();
  pdl_trigger();

  const int j_begin = (int)((long long)p.n_jobs * blockIdx.x / gridDim.x);
  const int j_end = (int)((long long)p.n_jobs * (blockIdx.x + 1) / gridDim.x);
  // job -> (head, first sample, samples): head-major
  auto job_head = [&](int job) { return job / p.groups; };

  if (warp == 8) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ph = 0;
      int res_head = -1;
      for (int job = j_begin; job < j_end; ++job, ph ^= 1u) {
        const int head = job_head(job), b0s = (job - head * p.groups) * p.spj;
        const int nseg = min(p.spj, p.Bp - b0s);
        const bool reuse = head == res_head;                         // prefix K rows stay; V is overwritten by dS at hd 128
        const bool reuse_v = reuse && HD != 128;
        const int col_h = head * HD;
        const int own_row0 = b0s * p.Ls;                              // first own row of the job (among the own rows)
        const int seg_boxes = nseg * (p.Lsp >> 4);
        // the prefix boxes run up to column na64: the rows between the prefix and the first own segment are masked, but the
        // tensor core multiplies them by the zeros of P / dS, so they must hold finite data, not stale shared memory
        const int k_boxes = (reuse ? 0 : (p.na64 >> 4)) + seg_boxes, v_boxes = (reuse_v ? 0 : (p.na64 >> 4)) + seg_boxes;
        mbar_wait(smem_free, ph ^ 1u, 600);                           // the previous job's dQ / dV / dK MMAs have retired
        if (job == j_begin + 1) TC_STAMP(1);
        mbar_arrive_expect_tx(ld_full, KB * (2 * Cfg::kSlabQ + (k_boxes + v_boxes) * 2048));
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(sQ + kb * Cfg::kSlabQ, &tmap_q, ld_full, col_h + kb * 64, p.Lc + own_row0, 0, kEvictFirst);
          tma_load_3d(sdO + kb * Cfg::kSlabQ, &tmap_do, ld_full, col_h + kb * 64, own_row0, 0, kEvictFirst);
        }
#pragma unroll
        for (int which = 0; which < 2; ++which) {                     // 0: K, 1: V
          const uint32_t slab = which ? sV : sK;
          const bool skip_prefix = which ? reuse_v : reuse;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t dst = slab + kb * Cfg::kSlabKV;
            const int col = (which + 1) * p.D + col_h + kb * 64;
            if (!skip_prefix) tma_load_rows(dst, &tmap_q, &tmap_kv64, &tmap_kv, ld_full, col, 0, 0, p.na64, kEvictLast);
            for (int s = 0; s < nseg; ++s)
              tma_load_rows(dst, &tmap_q, &tmap_kv64, &tmap_kv, ld_full, col, p.na64 + s * p.Lsp, p.Lc + own_row0 + s * p.Ls,
                            p.Lsp, kEvictNormal);
          }
        }
        if (job == j_begin + 1) TC_STAMP(2);
        res_head = head;
        if (job + 1 < j_end) {              // next job's Q / dO / own K / own V (and its prefix when the head changes) -> L2
          const int nh = job_head(job + 1), nb0 = (job + 1 - nh * p.groups) * p.spj;
          const int nns = min(p.spj, p.Bp - nb0), nrow0 = nb0 * p.Ls, ncol = nh * HD;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            tma_prefetch_3d(&tmap_q, ncol + kb * 64, p.Lc + nrow0, 0);
            tma_prefetch_3d(&tmap_do, ncol + kb * 64, nrow0, 0);
#pragma unroll
            for (int which = 1; which <= 2; ++which) {
              const int col = which * p.D + ncol + kb * 64;
              if (nh != head) tma_prefetch_rows(&tmap_q, &tmap_kv64, &tmap_kv, col, 0, p.na64);
              for (int s = 0; s < nns; ++s)
                tma_prefetch_rows(&tmap_q, &tmap_kv64, &tmap_kv, col, p.Lc + nrow0 + s * p.Ls, p.Lsp);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ph = 0;
      const uint32_t tS = tmem_base, tdP = tmem_base + 256, tdQ = tmem_base, tdV = tmem_base + 128, tdK = tmem_base + 256;
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, HD) | (1u << 16);                 // B MN-major
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, HD) | (1u << 15) | (1u << 16);    // A and B MN-major
      for (int job = j_begin; job < j_end; ++job, ph ^= 1u) {
        const int head = job_head(job), b0s = (job - head * p.groups) * p.spj;
        const int nseg = min(p.spj, p.Bp - b0s);
        const int nk = p.na64 + nseg * p.Lsp;                         // key columns (multiple of 16, <= 256)
        const uint32_t idesc_s = umma_idesc_bf16(128, (uint32_t)nk);
        mbar_wait(ld_full, ph, 610);
        if (job == j_begin + 1) TC_STAMP(3);
        mbar_wait(epi_done, ph ^ 1u, 611);                            // the previous job's gradients have left TMEM
        tc_fence_after();
        if (job == j_begin + 1) TC_STAMP(4);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t aq = umma_desc_sw128(sQ + kb * Cfg::kSlabQ), bk = umma_desc_sw128(sK + kb * Cfg::kSlabKV);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS, aq + 2u * k, bk + 2u * k, idesc_s, (kb | k) != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t ao = umma_desc_sw128(sdO + kb * Cfg::kSlabQ), bv = umma_desc_sw128(sV + kb * Cfg::kSlabKV);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tdP, ao + 2u * k, bv + 2u * k, idesc_s, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        if (job == j_begin + 1) TC_STAMP(5);
        mbar_wait(ds_full, ph, 612);                                  // P / dS are in shared memory, S / dP are consumed
        tc_fence_after();
        if (job == j_begin + 1) TC_STAMP(6);
        for (int ks = 0; ks < (nk >> 4); ++ks)                        // dQ = dS K
          umma_bf16(tdQ, umma_desc_sw128(sdS + (ks >> 2) * Cfg::kSlab) + 2u * (ks & 3),
                    umma_desc_mn_sw128(sK + ks * 2048, Cfg::kSlabKV), idesc_dq, ks != 0 ? 1u : 0u);
        const uint32_t ds_own = sdS + (p.na64 >> 6) * Cfg::kSlab;
#pragma unroll
        for (int kq = 0; kq < 8; ++kq) {                              // dV = P_own^T dO,  dK = dS_own^T Q   (K = 128 query rows)
          umma_bf16(tdV, umma_desc_mn_sw128(sP + kq * 2048, Cfg::kSlab), umma_desc_mn_sw128(sdO + kq * 2048, Cfg::kSlabQ),
                    idesc_kv, kq != 0 ? 1u : 0u);
          umma_bf16(tdK, umma_desc_mn_sw128(ds_own + kq * 2048, Cfg::kSlab), umma_desc_mn_sw128(sQ + kq * 2048, Cfg::kSlabQ),
                    idesc_kv, kq != 0 ? 1u : 0u);
        }
        umma_commit(smem_free);
        umma_commit(g_full);
        if (job == j_begin + 1) TC_STAMP(7);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ compute warps 0..7
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;                                // TMEM lane: query row (phase 1), own key (phase 2)
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int rx = r & 7;
    uint8_t* ds_row = base_ptr + Cfg::kOffdS + r * 128;
    uint8_t* p_row = base_ptr + Cfg::kOffP + r * 128;
    const uint8_t* do_row = base_ptr + Cfg::kOffdO + r * 128;
    const int r_div = r / p.Ls, m_div = r / p.Lsp;
    uint32_t ph = 0;
    for (int job = j_begin; job < j_end; ++job, ph ^= 1u) {
      const int head = job_head(job), b0s = (job - head * p.groups) * p.spj;
      const int nseg = min(p.spj, p.Bp - b0s);
      const int nk = p.na64 + nseg * p.Lsp;
      const int R = nseg * p.Ls;
      const int own_row0 = b0s * p.Ls;
      // ---- phase 1 role: query row r
      const bool row_valid = r < R;
      const int s = row_valid ? r_div : 0, t = r - s * p.Ls;
      const int own_lo = row_valid ? p.na64 + s * p.Lsp : (1 << 30);
      const int own_hi = row_valid ? own_lo + t : -1;
      const int n_prefix = row_valid ? p.Lc : 0;
      const float lse2 = row_valid ? p.lse_own[((int64_t)(b0s + s) * p.H + head) * p.Ls + t] * 1.4426950408889634f : INFINITY;
      // O row of this query (for delta), in flight before the operands land
      uint4 orow[HD / 8];
      {
        const uint4* op = reinterpret_cast<const uint4*>(p.out_own + (int64_t)(own_row0 + (row_valid ? r : 0)) * p.D +
                                                         (int64_t)head * HD);
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) orow[i] = __ldg(op + i);
      }
      uint32_t vmask = 0;                         // 32-column chunks in which some row of this quarter sees a key
      for (int c0 = 0, k = 0; c0 < nk; c0 += 32, ++k)
        if (c0 < p.Lc || __any_sync(0xffffffffu, c0 <= own_hi && c0 + 31 >= own_lo)) vmask |= 1u << k;
      if (job == j_begin && threadIdx.x == 0) TC_STAMP(0);
      if (job == j_begin + 1 && threadIdx.x == 0) TC_STAMP(8);
      if (lane == 0) mbar_wait(s_full, ph, 620);  // implies ld_full: Q / K / V / dO are in shared memory
      __syncwarp();
      tc_fence_after();
      if (job == j_begin + 1 && threadIdx.x == 0) TC_STAMP(9);
      // delta = rowsum(dO o O): dO from its swizzled tile (16-byte chunk q of row r sits at q ^ (r & 7))
      float delta = 0.0f;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 d = *reinterpret_cast<const uint4*>(do_row + kb * Cfg::kSlabQ + ((q ^ rx) << 4));
          const uint4 o = orow[kb * 8 + q];
          const uint32_t dw[4] = {d.x, d.y, d.z, d.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) delta += bf16_lo(dw[i]) * bf16_lo(ow[i]) + bf16_hi(dw[i]) * bf16_hi(ow[i]);
        }
      }
      if (!row_valid) delta = 0.0f;
      if (p.delta_own != nullptr && half == 0 && row_valid)
        p.delta_own[((int64_t)(b0s + s) * p.H + head) * p.Ls + t] = delta;
      // the two warps of a quarter alternate chunks; chunks nobody sees are zero-filled
      for (int c0 = 32 * half, k = half; c0 < nk; c0 += 64, k += 2) {
        uint8_t* dsd = ds_row + (c0 >> 6) * Cfg::kSlab;
        const int q0 = (c0 & 63) >> 3;
        const bool own_chunk = c0 >= p.na64;
        uint8_t* pd = p_row + (own_chunk ? ((c0 - p.na64) >> 6) : 0) * Cfg::kSlab;
        if (!(vmask & (1u << k))) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            *reinterpret_cast<uint4*>(dsd + (((q0 + g) ^ rx) << 4)) = make_uint4(0, 0, 0, 0);
            if (own_chunk) *reinterpret_cast<uint4*>(pd + (((q0 + g) ^ rx) << 4)) = make_uint4(0, 0, 0, 0);
          }
          continue;
        }
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tq + c0, sv);
        tmem_ld_32x32(tq + 256 + c0, dv);
        tmem_ld_wait();
        const uint32_t bits = chunk_mask(c0, n_prefix, own_lo, own_hi);
        float pv[32], dsv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool vis = (bits & (1u << j)) != 0;               // selects, not products: a masked column may hold anything
          const float e = ex2_approx(fmaf(__uint_as_float(sv[j]), p.scale_log2e, -lse2));
          pv[j] = vis ? e : 0.0f;
          dsv[j] = vis ? e * (__uint_as_float(dv[j]) - delta) * p.scale : 0.0f;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<uint4*>(dsd + (((q0 + g) ^ rx) << 4)) =
              make_uint4(pack_bf16(dsv[8 * g], dsv[8 * g + 1]), pack_bf16(dsv[8 * g + 2], dsv[8 * g + 3]),
                         pack_bf16(dsv[8 * g + 4], dsv[8 * g + 5]), pack_bf16(dsv[8 * g + 6], dsv[8 * g + 7]));
          if (own_chunk)
            *reinterpret_cast<uint4*>(pd + (((q0 + g) ^ rx) << 4)) =
                make_uint4(pack_bf16(pv[8 * g], pv[8 * g + 1]), pack_bf16(pv[8 * g + 2], pv[8 * g + 3]),
                           pack_bf16(pv[8 * g + 4], pv[8 * g + 5]), pack_bf16(pv[8 * g + 6], pv[8 * g + 7]));
        }
      }
      // own columns past the job's segments (last, smaller group): P must still be zero there — those are M rows of dV
      for (int c0 = ((nk + 31) & ~31) + 32 * half; c0 < p.na64 + 128; c0 += 64) {
        uint8_t* pd = p_row + ((c0 - p.na64) >> 6) * Cfg::kSlab;
        uint8_t* dsd = ds_row + (c0 >> 6) * Cfg::kSlab;
        const int q0 = (c0 & 63) >> 3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<uint4*>(pd + (((q0 + g) ^ rx) << 4)) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(dsd + (((q0 + g) ^ rx) << 4)) = make_uint4(0, 0, 0, 0);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_full);
      if (job == j_begin + 1 && lane == 0) TC_STAMP(16 + warp);

      // ---- phase 2 role: lane r = query row for dQ, own key index for dV / dK
      const int ms = m_div, mj = r - ms * p.Lsp;                       // own key r: sample ms of the job, token mj
      const bool key_valid = ms < nseg && mj < p.Ls;
      __nv_bfloat16* dq_dst = p.dqkv_own + (int64_t)(own_row0 + r) * 3 * p.D + (int64_t)head * HD;
      __nv_bfloat16* dk_dst = p.dqkv_own + (int64_t)(own_row0 + ms * p.Ls + mj) * 3 * p.D + p.D + (int64_t)head * HD;
      __nv_bfloat16* dv_dst = dk_dst + p.D;
      // table rows of lanes without a query / key are never used: point them at row Lc so that the (unconditional,
      // early) loads below stay inside the [Lc + Ls, hd/2] tables
      const int tq_pos = p.Lc + (row_valid ? t : 0), tk_pos = p.Lc + (key_valid ? mj : 0);
      const float* cq = p.rope_cos ? p.rope_cos + (int64_t)tq_pos * (HD / 2) : nullptr;
      const float* sq = p.rope_sin ? p.rope_sin + (int64_t)tq_pos * (HD / 2) : nullptr;
      const float* ck = p.rope_cos ? p.rope_cos + (int64_t)tk_pos * (HD / 2) : nullptr;
      const float* sk = p.rope_sin ? p.rope_sin + (int64_t)tk_pos * (HD / 2) : nullptr;
      // RoPE table rows of this lane's chunk pair, in flight before the gradients are ready.  With Ls a multiple of 16 lane r
      // is query (s, t) AND own key (s, t): one pair of table rows serves dQ and dK.
      constexpr int kHalfChunks = HD / 64;
      const int c_pair = (HD == 128) ? half : 0;
      const bool same_rows = p.Lsp == p.Ls;
      float cs[32], sn[32];
      auto load_tables = [&](const float* cosr, const float* sinr) {
        const float4* c4 = reinterpret_cast<const float4*>(cosr + 32 * c_pair);
        const float4* s4 = reinterpret_cast<const float4*>(sinr + 32 * c_pair);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 cv = __ldg(c4 + j), sv4 = __ldg(s4 + j);
          cs[4 * j] = cv.x; cs[4 * j + 1] = cv.y; cs[4 * j + 2] = cv.z; cs[4 * j + 3] = cv.w;
          sn[4 * j] = sv4.x; sn[4 * j + 1] = sv4.y; sn[4 * j + 2] = sv4.z; sn[4 * j + 3] = sv4.w;
        }
      };
      const bool roped = p.rope_cos != nullptr;
      const bool first_is_q = (HD == 128) || half == 0;               // hd 64: half 1 only handles dK
      if (roped) {
        if (first_is_q) load_tables(cq, sq); else load_tables(ck, sk);
      }
      if (lane == 0) mbar_wait(g_full, ph, 621);
      __syncwarp();
      tc_fence_after();
      if (job == j_begin + 1 && threadIdx.x == 0) TC_STAMP(10);
      // a (lo, hi) chunk pair of a rotated gradient: rotate back with the loaded table rows and store;
      // lo chunk c holds columns [32c, 32c + 32)
      auto store_pair = [&](uint32_t tcol, __nv_bfloat16* dst, bool valid) {
        uint32_t lo[32], hi[32];
        tmem_ld_32x32(tq + tcol + 32 * c_pair, lo);
        tmem_ld_32x32(tq + tcol + 32 * (c_pair + kHalfChunks), hi);
        tmem_ld_wait();
        if (!valid) return;
        float a[32], b[32];
        if (roped) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(lo[j]), y = __uint_as_float(hi[j]);
            a[j] = x * cs[j] + y * sn[j];
            b[j] = y * cs[j] - x * sn[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(lo[j]); b[j] = __uint_as_float(hi[j]); }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {           // 32-byte stores: one full sector per lane
          st_global_v8(dst + 32 * c_pair + 16 * g, pack_bf16(a[16 * g], a[16 * g + 1]), pack_bf16(a[16 * g + 2], a[16 * g + 3]),
                       pack_bf16(a[16 * g + 4], a[16 * g + 5]), pack_bf16(a[16 * g + 6], a[16 * g + 7]),
                       pack_bf16(a[16 * g + 8], a[16 * g + 9]), pack_bf16(a[16 * g + 10], a[16 * g + 11]),
                       pack_bf16(a[16 * g + 12], a[16 * g + 13]), pack_bf16(a[16 * g + 14], a[16 * g + 15]));
          st_global_v8(dst + 32 * (c_pair + kHalfChunks) + 16 * g, pack_bf16(b[16 * g], b[16 * g + 1]),
                       pack_bf16(b[16 * g + 2], b[16 * g + 3]), pack_bf16(b[16 * g + 4], b[16 * g + 5]),
                       pack_bf16(b[16 * g + 6], b[16 * g + 7]), pack_bf16(b[16 * g + 8], b[16 * g + 9]),
                       pack_bf16(b[16 * g + 10], b[16 * g + 11]), pack_bf16(b[16 * g + 12], b[16 * g + 13]),
                       pack_bf16(b[16 * g + 14], b[16 * g + 15]));
        }
      };
      if constexpr (HD == 128) {
        // half 0: dQ pair (0, 2), dK pair (0, 2), dV chunks 0, 1;   half 1: dQ pair (1, 3), dK pair (1, 3), dV chunks 2, 3
        store_pair(0, dq_dst, row_valid);
        if (roped && !same_rows) load_tables(ck, sk);
        store_pair(256, dk_dst, key_valid);
        // dV: two adjacent chunks, no rotation (store_pair with a pair distance of one chunk would need another shape)
        {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(tq + 128 + 64 * half, v0);
          tmem_ld_32x32(tq + 128 + 64 * half + 32, v1);
          tmem_ld_wait();
          if (key_valid) {
            auto store32 = [&](__nv_bfloat16* dst, const uint32_t (&v)[32]) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                uint32_t w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  w[q] = pack_bf16(__uint_as_float(v[16 * g + 2 * q]), __uint_as_float(v[16 * g + 2 * q + 1]));
                st_global_v8(dst + 16 * g, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
              }
            };
            store32(dv_dst + 64 * half, v0);
            store32(dv_dst + 64 * half + 32, v1);
          }
        }
      } else {
        // hd 64: half 0: dQ pair (0, 1) + dV chunk 0;   half 1: dK pair (0, 1) + dV chunk 1
        if (half == 0) store_pair(0, dq_dst, row_valid);
        else store_pair(256, dk_dst, key_valid);
        uint32_t v0[32];
        tmem_ld_32x32(tq + 128 + 32 * half, v0);
        tmem_ld_wait();
        if (key_valid) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              w[q] = pack_bf16(__uint_as_float(v0[16 * g + 2 * q]), __uint_as_float(v0[16 * g + 2 * q + 1]));
            st_global_v8(dv_dst + 32 * half + 16 * g, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(epi_done);
      if (job == j_begin + 1 && lane == 0) TC_STAMP(24 + warp);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// 0 off | 1 auto (default) | 2 whenever the shape fits (tests) — MTS_ATTN_TC / mts_set_option("attn_tc", v)
static int g_attn_tc = -1;
static int attn_tc_mode() {
  if (g_attn_tc < 0) {
    const char* e = getenv("MTS_ATTN_TC");
    g_attn_tc = e ? atoi(e) : 1;
    if (g_attn_tc < 0 || g_attn_tc > 2) g_attn_tc = 1;
  }
  return g_attn_tc;
}
void attn_tc_set(int v) { g_attn_tc = (v < 0 || v > 2) ? 1 : v; }

// Whether the tensor-memory kernel takes this shape.  The answer depends on (Bp, H, L, hd) only — never on how much of L
// is a shared prefix — so the shared-prefix and the per-sample layouts of one model take the same route, which is what
// keeps their outputs bit-identical.  Head dim 64 with short sequences or little work (one job per SM or less: launch +
// pipeline-fill bound) stays on the mma.sync kernels, which start faster (tools/bench_attn.py --hd64,
// profiles/r02_attn_tc.md).
bool attn_tc_eligible(int L, int Lc, int hd, int Bp, int H) {
  const int mode = attn_tc_mode();
  if (mode == 0 || !(hd == 64 || hd == 128) || Lc < 0 || Lc >= L) return false;
  // L <= 240 always fits 256 key columns; up to 256 when no dummy columns are needed (prefix a multiple of 16)
  if (!(L <= 240 || (L <= 256 && (Lc % 16) == 0))) return false;
  if (mode == 1 && hd == 64 && (L < 160 || (long long)Bp * H * L < 65536)) return false;
  return true;
}

template <int HD>
static int launch_attn_tc_t(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int L, int Lc, int H, float scale,
                            cudaStream_t stream) {
  using Cfg = AttnTcCfg<HD>;
  auto kern = attn_fwd_tc_kernel<HD>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn_fwd_tc_kernel)", e);
    attr_done = true;
  }
  if (reinterpret_cast<uintptr_t>(out) & 31)
    return set_error(MTS_ERR_INVALID_ARG, "attention: the output must be 32-byte aligned (32-byte stores)");
  AttnTcParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.Bp = Bp; p.Lc = Lc; p.Ls = L - Lc; p.H = H; p.D = H * HD;
  p.n_pt = (Lc + 127) / 128;
  p.n_qt = (p.Ls + 127) / 128;
  const int lsp = ((Lc & 15) + p.Ls + 15) & ~15;
  int sample_jobs;
  if (p.n_qt == 1) {
    int spj = 128 / p.Ls;
    const int by_cols = (256 - ((Lc + 15) & ~15)) / lsp;
    if (by_cols < spj) spj = by_cols;
    if (spj > Bp) spj = Bp;
    if (spj < 1) spj = 1;
    p.spj = spj;
    sample_jobs = (Bp + spj - 1) / spj;
  } else {
    p.spj = 1;
    sample_jobs = Bp * p.n_qt;
  }
  p.jobs_per_head = p.n_pt + sample_jobs;
  const int64_t n_jobs = (int64_t)p.jobs_per_head * H;
  if (n_jobs > 0x7fffffffLL) return set_error(MTS_ERR_INVALID_ARG, "attention: too many tiles");
  p.n_jobs = (int)n_jobs;
  p.scale_log2e = scale * 1.4426950408889634f;
  const int64_t rows = (int64_t)Lc + (int64_t)Bp * p.Ls;
  CUtensorMap tq, tkv, tkv64;
  int rc = get_tmap_3d(&tq, qkv, 3 * (int64_t)p.D, rows, 1, 3 * (int64_t)p.D, rows * 3 * p.D, 64, 128, 2);
  if (rc) return rc;
  rc = get_tmap_3d(&tkv, qkv, 3 * (int64_t)p.D, rows, 1, 3 * (int64_t)p.D, rows * 3 * p.D, 64, 16, 2);
  if (rc) return rc;
  rc = get_tmap_3d(&tkv64, qkv, 3 * (int64_t)p.D, rows, 1, 3 * (int64_t)p.D, rows * 3 * p.D, 64, 64, 2);
  if (rc) return rc;
  const int grid = p.n_jobs < num_sms() ? p.n_jobs : num_sms();
  static long long* dbg_buf = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) { const char* e = getenv("MTS_ATTN_TC_DBG"); dbg_on = (e && e[0] == '1') ? 1 : 0; }
  p.dbg = nullptr;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 24 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 24 * sizeof(long long), stream);
    p.dbg = dbg_buf;
  }
  LAUNCH_PDL(kern, grid, kTcThreads, Cfg::kSmemBytes, stream, tq, tkv, tkv64, p);
  count_launch();
  if (dbg_on) {
    long long h[24];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    static const char* names[14] = {"start", "K issued", "V issued", "q_full seen", "S issued", "p_full+v_full seen", "O issued",
                                    "s_full seen", "pass1 done", "P written", "o_full seen", "job0 done", "last job done", "end"};
    fprintf(stderr, "[attn_tc dbg] jobs=%d grid=%d spj=%d:", p.n_jobs, grid, p.spj);
    for (int i = 1; i < 14; ++i) fprintf(stderr, " %s=%lld", names[i], h[i] ? h[i] - h[0] : -1);
    for (int i = 16; i < 20; ++i) fprintf(stderr, " P of warp %d=%lld", i & 3, h[i] ? h[i] - h[0] : -1);
    fprintf(stderr, "\n");
  }
  return check_launch("attn_fwd_tc_kernel");
}

int launch_attn_tc(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int L, int Lc, int H, int hd, float scale,
                   cudaStream_t stream) {
  if (hd == 128) return launch_attn_tc_t<128>(qkv, out, lse, Bp, L, Lc, H, scale, stream);
  return launch_attn_tc_t<64>(qkv, out, lse, Bp, L, Lc, H, scale, stream);
}

// The tensor-memory backward covers the frozen-backbone case (own rows only) when a sample's own rows fit one tile and
// prefix + own columns fit the 256-column score tile with the own columns starting on a multiple of 64.
bool attn_tc_bwd_eligible(int Lc, int Ls, int hd, int Bp, int H, const float* rc, const float* rs) {
  const int mode = attn_tc_mode();
  if ((reinterpret_cast<uintptr_t>(rc) & 15) || (reinterpret_cast<uintptr_t>(rs) & 15)) return false;   // float4 table loads
  if (mode == 0 || !(hd == 64 || hd == 128) || Lc < 16 || Ls < 1 || Ls > 128) return false;
  const int na64 = (Lc + 63) & ~63, lsp = (Ls + 15) & ~15;
  if (na64 > 128 || na64 + lsp > 256) return false;
  if (mode == 1 && hd == 64 && (long long)Bp * H * (Lc + Ls) < 65536) return false;
  return true;
}

template <int HD>
static int launch_attn_bwd_tc_t(const uint16_t* qkv, const float* rc, const float* rs, const uint16_t* out_own,
                                const uint16_t* dout_own, const float* lse_own, float* delta_own, uint16_t* dqkv_own, int Bp,
                                int Lc, int Ls, int H, float scale, cudaStream_t stream) {
  using Cfg = AttnTcBwdCfg<HD>;
  auto kern = attn_bwd_tc_kernel<HD>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn_bwd_tc_kernel)", e);
    attr_done = true;
  }
  if ((reinterpret_cast<uintptr_t>(dqkv_own) & 31) || (reinterpret_cast<uintptr_t>(out_own) & 15))
    return set_error(MTS_ERR_INVALID_ARG, "attention backward: dqkv must be 32-byte and out 16-byte aligned (vector accesses)");
  AttnTcBwdParams p;
  p.out_own = reinterpret_cast<const __nv_bfloat16*>(out_own);
  p.lse_own = lse_own;
  p.rope_cos = rc; p.rope_sin = rs;
  p.dqkv_own = reinterpret_cast<__nv_bfloat16*>(dqkv_own);
  p.delta_own = delta_own;
  p.Bp = Bp; p.Lc = Lc; p.Ls = Ls; p.H = H; p.D = H * HD;
  p.na64 = (Lc + 63) & ~63;
  p.Lsp = (Ls + 15) & ~15;
  int spj = 128 / Ls;
  const int by_cols = (256 - p.na64) / p.Lsp, by_m = 128 / p.Lsp;    // score columns; the 128-row M block of dK / dV
  if (by_cols < spj) spj = by_cols;
  if (by_m < spj) spj = by_m;
  if (spj > Bp) spj = Bp;
  if (spj < 1) spj = 1;
  p.spj = spj;
  p.groups = (Bp + spj - 1) / spj;
  const int64_t n_jobs = (int64_t)p.groups * H;
  if (n_jobs > 0x7fffffffLL) return set_error(MTS_ERR_INVALID_ARG, "attention backward: too many tiles");
  p.n_jobs = (int)n_jobs;
  p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  const int64_t rows = (int64_t)Lc + (int64_t)Bp * Ls, own_rows = (int64_t)Bp * Ls;
  CUtensorMap tq, tkv, tkv64, tdo;
  int rc_ = get_tmap_3d(&tq, qkv, 3 * (int64_t)p.D, rows, 1, 3 * (int64_t)p.D, rows * 3 * p.D, 64, 128, 2);
  if (rc_) return rc_;
  rc_ = get_tmap_3d(&tkv, qkv, 3 * (int64_t)p.D, rows, 1, 3 * (int64_t)p.D, rows * 3 * p.D, 64, 16, 2);
  if (rc_) return rc_;
  rc_ = get_tmap_3d(&tkv64, qkv, 3 * (int64_t)p.D, rows, 1, 3 * (int64_t)p.D, rows * 3 * p.D, 64, 64, 2);
  if (rc_) return rc_;
  rc_ = get_tmap_3d(&tdo, dout_own, (int64_t)p.D, own_rows, 1, (int64_t)p.D, own_rows * p.D, 64, 128, 2);
  if (rc_) return rc_;
  const int grid = p.n_jobs < num_sms() ? p.n_jobs : num_sms();
  static long long* dbg_buf = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) { const char* e = getenv("MTS_ATTN_TC_DBG"); dbg_on = (e && e[0] == '1') ? 1 : 0; }
  p.dbg = nullptr;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 32 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 32 * sizeof(long long), stream);
    p.dbg = dbg_buf;
  }
  LAUNCH_PDL(kern, grid, kTcBwdThreads, Cfg::kSmemBytes, stream, tq, tkv, tkv64, tdo, p);
  count_launch();
  if (dbg_on) {
    long long h[32];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    static const char* names[11] = {"", "producer: smem free", "loads issued", "MMA: operands landed", "TMEM free", "S + dP issued",
                                    "dS seen", "dQ dV dK issued", "compute: waits for S", "S seen", "gradients seen"};
    fprintf(stderr, "[attn_bwd_tc dbg] jobs=%d grid=%d spj=%d, second job of block 0 (cycles since its first job's S wait):", p.n_jobs, grid, p.spj);
    for (int i = 1; i < 11; ++i) fprintf(stderr, " %s=%lld", names[i], h[i] ? h[i] - h[0] : -1);
    for (int i = 16; i < 24; ++i) fprintf(stderr, " dS of warp %d=%lld", i - 16, h[i] ? h[i] - h[0] : -1);
    for (int i = 24; i < 32; ++i) fprintf(stderr, " done warp %d=%lld", i - 24, h[i] ? h[i] - h[0] : -1);
    fprintf(stderr, "\n");
  }
  return check_launch("attn_bwd_tc_kernel");
}

int launch_attn_bwd_tc(const uint16_t* qkv, const float* rc, const float* rs, const uint16_t* out_own, const uint16_t* dout_own,
                       const float* lse_own, float* delta_own, uint16_t* dqkv_own, int Bp, int Lc, int Ls, int H, int hd,
                       float scale, cudaStream_t stream) {
  if (hd == 128)
    return launch_attn_bwd_tc_t<128>(qkv, rc, rs, out_own, dout_own, lse_own, delta_own, dqkv_own, Bp, Lc, Ls, H, scale, stream);
  return launch_attn_bwd_tc_t<64>(qkv, rc, rs, out_own, dout_own, lse_own, delta_own, dqkv_own, Bp, Lc, Ls, H, scale, stream);
}

}  // namespace mts
