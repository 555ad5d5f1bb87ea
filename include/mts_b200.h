/*
 * mts_b200.h — C ABI of libmtsb200.so, the sm_100a kernel stack behind MedTsLLM's hot path.
 *
 * This is the drop-in boundary.  The reference (flixpar/med-ts-llm) is pure Python
 * and reaches its "kernels" through PyTorch/HuggingFace library calls; every entry point below
 * replaces one of those call sites and cites it as   ref: <file>:<line>   (paths relative to the
 * reference tree; `HF:` = site-packages/transformers 5.5.0, the reference's third-party backbone).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Every pointer is a DEVICE pointer unless its
 *     name starts with `h_`.  The library owns no tensor memory: callers allocate everything.
 *   - All launches are asynchronous on the cudaStream_t passed as `mts_stream_t`; no entry point
 *     synchronises or allocates device memory.  (Host-side caches: TMA descriptors only.)
 *   - Return value: MTS_OK (0) or an mts_status error code; `mts_last_error()` returns a
 *     thread-local human-readable message for the last failure on the calling thread.
 *   - Row-major everywhere.  "ld*" strides and batch strides are in ELEMENTS, not bytes.
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t).  Residual stream and statistics are fp32.
 *   - There is NO CPU fallback.  On a machine without an sm_100 GPU every compute entry point
 *     returns MTS_ERR_CUDA.
 */
#ifndef MTS_B200_H_
#define MTS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTS_ABI_VERSION 2

typedef void* mts_stream_t; /* cudaStream_t */

typedef enum mts_status {
  MTS_OK = 0,
  MTS_ERR_INVALID_ARG = 1,
  MTS_ERR_CUDA = 2,
  MTS_ERR_UNSUPPORTED = 3
} mts_status;

typedef enum mts_dtype { MTS_BF16 = 0, MTS_F32 = 1 } mts_dtype;

/* ------------------------------------------------------------------------------------------ */
/* Library                                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* ABI version (MTS_ABI_VERSION of the build). */
int mts_version(void);
/* Thread-local message for the last non-OK status returned on this thread ("" if none). */
const char* mts_last_error(void);
/* Number of kernels this library has launched since load (process-wide; bench `gpu_launches`). */
int64_t mts_launch_count(void);
/* Drop the host-side TMA descriptor cache. */
int mts_clear_caches(void);
/* Process-wide switches.
 *   "gemm_2cta" (0/1; default 1, env MTS_GEMM_2CTA=0): route 256-wide GEMM tiles to the CTA-pair kernel
 *               (tcgen05 cta_group::2, 256x256 tile per two SMs) when the cost model prefers it.
 *   "gemm_force" (0 auto | 1 single-CTA kernel | 2 CTA-pair kernel; default 0): override that cost model (experiments,
 *               tools/bench_gemm.py --force-sweep).
 *   "streamk"   (0 off | 1 auto | 2 force | 3 even split-K: tiles x s CTAs, s <= 4; default 0, env MTS_STREAMK): see
 *               mts_gemm_args.sk_workspace.
 *   "gemm_ksplit" (-1 auto | 0 off | 2 / 4 forced whenever legal; default -1, env MTS_GEMM_KSPLIT): cluster split-K of the
 *               single-CTA kernel — clusters of s CTAs share one tile, each runs 1/s of the k loop, the partial column
 *               parts are exchanged through distributed shared memory and every CTA finishes one part (STORE /
 *               RESID_ADD / GELU_NEW epilogues, batch 1, bf16 operands).  Auto: pairs on 128-wide tiles when
 *               tiles x 2 <= SMs and k >= 2048.  Regroups the fp32 k-sum (deterministic, but not the order of the
 *               unsplit schedule).
 *   "attn_tc"   (0 off | 1 auto | 2 whenever the shape fits; default 1, env MTS_ATTN_TC): the causal-attention forward on
 *               tcgen05 / tensor memory (sequences of at most 256 positions, head dim 64 / 128) instead of the mma.sync
 *               kernels.  "auto" decides from (samples, heads, positions, head dim) only, so that the shared-prefix and
 *               the per-sample layout of one model always run the same arithmetic.
 *   "epi_direct" (0/1; default 1, env MTS_EPI_DIRECT=0): GEMM epilogues store full 32-column chunks straight from the
 *               registers with 32-byte accesses when the rows of D are 32-byte aligned (0 = always stage through smem).
 *   "pdl"       (0/1; default 1, env MTS_PDL=0): launch the per-layer kernels with programmatic stream serialization
 *               (their prologues overlap the previous kernel's tail; they block in griddepcontrol.wait before
 *               touching its results).
 * Environment only: MTS_ATTN_FUSED_BWD=0 keeps the separate delta / dQ / dK,dV attention-backward kernels. */
int mts_set_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------ */
/* K1+K2  RevIN + patching + TokenEmbedding (fused front end)                                  */
/* ------------------------------------------------------------------------------------------ */
/*
 * ref: models/layers/RevIN.py:37-56 (statistics + normalise), models/layers/embed.py:155-163
 *      (ReplicationPad1d), :186-197 (PatchEmbedding.forward: unfold(P,S)), :29-46 (TokenEmbedding:
 *      Conv1d(P->d_model, k=3, circular over the patch axis, no bias)), and the `concat` reshape at
 *      models/medtsllm.py:276-279.
 *
 *   x       [B,T,C] fp32
 *   w_conv  [d_model, P, 3] fp32     (patch_embedding.value_embedding.tokenConv.weight)
 *   mean, stdev  [B,C] fp32 out      (RevIN statistics; stdev = sqrt(var_biased + eps))
 *   out     bf16 and/or fp32, either pointer may be NULL:
 *             concat_layout = 1 :  [B, N, C*d_model]   (feature-major inside a token)
 *             concat_layout = 0 :  [B*C, N, d_model]
 *   N = (T + S - P)/S + 1 patches; patch n covers padded samples n*S .. n*S+P-1, where the padded
 *   series repeats sample T-1 S times at the end.
 */
int mts_revin_patch_embed(const float* x, const float* w_conv, float* mean, float* stdev,
                          uint16_t* out_bf16, float* out_f32, int B, int T, int C, int P, int S,
                          int d_model, int concat_layout, float eps, mts_stream_t stream);

/*
 * Pure index work of the same path: gathers the (un-normalised) patches so that the patch index
 * map can be checked bit-exactly against `unfold` (ref: models/layers/embed.py:188-190).
 *   patches [B*C, N, P] fp32 out;  patches[(b*C+c), n, p] = x[b, min(n*S+p, T-1), c]
 */
int mts_patch_gather(const float* x, float* patches, int B, int T, int C, int P, int S,
                     mts_stream_t stream);

/*
 * Backward of the front end w.r.t. the conv weight (the only trainable tensor in it; the RevIN
 * statistics are detached in the reference, RevIN.py:42-43).
 *   dout  [same layout as out] fp32,  dw_conv [d_model,P,3] fp32 (overwritten)
 */
int mts_revin_patch_embed_bwd(const float* x, const float* mean, const float* stdev,
                              const float* dout, float* dw_conv, int B, int T, int C, int P, int S,
                              int d_model, int concat_layout, mts_stream_t stream);

/* RevIN "denorm": y[b,t,c] = y[b,t,c]*stdev[b,c] + mean[b,c]   (ref: RevIN.py:58-69) */
int mts_revin_denorm(float* y, const float* mean, const float* stdev, int B, int T, int C,
                     mts_stream_t stream);

/* GPT4TS front end (BASELINE configs[0]; ref: models/gpt4ts.py:126-138, :151-164, :200-212, :230-242 and
 * models/layers/embed.py:8-46, 109-131 with x_mark = None): per-(sample, channel) normalisation over time
 * (mean, sqrt(biased var + eps)), the 3-tap circular TokenEmbedding conv over TIME with the C variables as
 * in-channels, plus the fixed sinusoid table pe [T, d_model].
 *   x fp32 [B,T,C]; w_conv fp32 [d_model, C, 3]; mean, stdev fp32 [B, C] out
 *   mode 0: out_t bf16 [B, d_model, ld_t] (transposed; columns T..ld_t-1 zero) — B operand of the time-axis Linear
 *   mode 1: out_x fp32 [B, T, D] = embedding, zero-padded to D columns (:212), + wpe[t] (HF GPT-2 position table)
 * (anomaly_detection needs no front end: its per-time-step "segments" centre the series to exactly zero, :155-164.) */
int mts_gpt4ts_embed(const float* x, const float* w_conv, const float* pe, const float* wpe, float* mean,
                     float* stdev, uint16_t* out_t, float* out_x, uint16_t* out_nt, int B, int T, int C, int d_model,
                     int D, int ld_t, int mode, float eps, mts_stream_t stream);
/* (mode 0 only, optional) out_nt bf16 [B, T, d_model]: the same embedding, not transposed — the B operand of the
 * time-axis Linear's weight gradient in training. */
/* Gradient of the TokenEmbedding conv weight (autograd of models/layers/embed.py:29-46 as GPT4TS calls it):
 * dw[d, c, k] = sum_{b,t} denc(b, d, t) * xn[b, (t+k-1) mod T, c];  denc fp32, element (b, d, t) at b*sb + d*sd + t*st
 * ([B, d_model, T] after the time-axis Linear, [B, T, D] straight from the residual stream); dw fp32 [d_model, C, 3]. */
int mts_gpt4ts_conv_wgrad(const float* x, const float* mean, const float* stdev, const float* denc, float* dw,
                          int B, int T, int C, int d_model, int64_t sb, int64_t sd, int64_t st, mts_stream_t stream);
/* LayerNorm (layernorm != 0) / RMSNorm parameter gradients, stage 1: partial fp32 [ceil(rows/32), 2*D] =
 * per-32-row sums of dy * xhat (first D columns) and dy (last D columns); sum the partials with mts_colsum.
 * GPT4TS trains the GPT-2 LayerNorms (models/gpt4ts.py:47-53).  x fp32 [rows, ldx], dy bf16 [rows, D]. */
int mts_norm_wgrad_partial(const float* x, int64_t ldx, const uint16_t* dy, float* partial, int rows, int D,
                           float eps, int layernorm, mts_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K3/K4/K7/K9/K10/K11/K12/K13  tcgen05 GEMM with fused epilogues                              */
/* ------------------------------------------------------------------------------------------ */
/*
 * D[b] = epilogue( alpha * A[b] (m x k) * B[b]^T (k x n) + bias )       ("NT": both K-major)
 *
 * Replaces every nn.Linear / Conv1D / einsum contraction on the path:
 *   ref: models/medtsllm.py:281 (mapping_layer), :566-591 (ReprogrammingLayer projections and the
 *        two einsums), :358 (embedding_downsample_layer), :541-552 (FlattenHead.linear);
 *        HF:models/llama/modeling_llama.py:182-184,262-264,288; HF:models/gpt2/modeling_gpt2.py:185,
 *        223,238-243 + HF:pytorch_utils.py:119-123 (Conv1D = addmm).
 *
 * A, B: bf16.  TMA-staged 128B-swizzled shared-memory tiles (BLOCK_M=128, BLOCK_K=64, BLOCK_N in
 * {64,128,256}), tcgen05.mma kind::f16 into fp32 TMEM accumulators (double-buffered), persistent
 * over min(tiles, SMs) CTAs.  Requirements: lda, ldb multiples of 8 (k itself may be anything: the
 * TMA zero-fills past k); a, b 16-byte aligned; k >= 1; batch strides multiples of 8 (or 0 = the
 * operand is shared by all batches).  D rows that are 16-byte aligned get vector stores, others a
 * scalar epilogue.
 */
typedef enum mts_epilogue {
  MTS_EPI_STORE = 0,     /* D = v                      (D bf16 or fp32)                          */
  MTS_EPI_RESID_ADD = 1, /* D = C + v   (D, C fp32; C = D when args.c is NULL) or, with bf16 D,   */
                         /* D = bf16(D + v) in place (LoRA side GEMMs accumulating into q / v)   */
  MTS_EPI_GELU_NEW = 2,  /* D = gelu_new(v)            (D bf16; HF:activations.py:59-66)         */
  MTS_EPI_SWIGLU = 3,    /* D[:, j] = silu(v_gate[j]) * v_up[j]   (D bf16, n/2 columns).  B rows  */
                         /* must be packed by mts_pack_gate_up: blocks of 128 gate rows followed  */
                         /* by the matching 128 up rows (HF:models/llama/modeling_llama.py:182-184) */
  MTS_EPI_ROPE_QK = 4    /* fused qkv projection with rotate-half RoPE applied (in fp32, from the   */
                         /* accumulators) to output columns < rope_cols; D bf16; head dim 64 / 128, */
                         /* position = row % rope_L (HF:models/llama/modeling_llama.py:139-168);    */
                         /* see rope_prefix for the shared-prefix row layout                         */
} mts_epilogue;

typedef enum mts_bias_axis { MTS_BIAS_NONE = 0, MTS_BIAS_N = 1, MTS_BIAS_M = 2 } mts_bias_axis;

typedef struct mts_gemm_args {
  const void* a;     /* bf16 [batch][m][k]                                                        */
  const void* b;     /* bf16 [batch][n][k]                                                        */
  void* d;           /* [batch][m][n'] (n' = n, or n/2 for SWIGLU); or [batch][n][m] if d_transposed */
  const float* bias; /* fp32 [n] (MTS_BIAS_N) or [m] (MTS_BIAS_M) or NULL                         */
  const float* c;    /* RESID_ADD only: D = C + v with C laid out like D; NULL = in place (C = D)  */
  int64_t lda, ldb, ldd;
  int64_t a_batch_stride, b_batch_stride, d_batch_stride;
  int32_t m, n, k, batch;
  int32_t d_dtype;      /* mts_dtype */
  int32_t epilogue;     /* mts_epilogue */
  int32_t bias_axis;    /* mts_bias_axis */
  int32_t d_transposed; /* STORE only: element (row i, col j) goes to d[j*ldd + i]               */
  int32_t block_n;      /* 0 = choose; else 64 / 128 / 256                                        */
  float alpha;
  /* MTS_EPI_ROPE_QK only: fp32 tables [>= rope_L, rope_hd/2], sequence length, head dim, #rotated columns */
  const float* rope_cos;
  const float* rope_sin;
  int32_t rope_L, rope_hd, rope_cols;
  int32_t rope_prefix; /* shared-prefix row layout: position = row (row < rope_prefix), else rope_prefix + (row - rope_prefix) % rope_L */
  /* MTS_EPI_SWIGLU only, optional: bf16 [m, ld_aux >= n] copy of the gate/up pre-activations (same packed   */
  /* column order as the weight rows) kept for the backward; batch must be 1 when used                       */
  void* aux;
  int64_t ld_aux;
  /* ABI 2 — evaluation parity mode (the reference evaluates with fp32 weights and TF32 matmuls, tasks/base.py:19-22):
   * ab_dtype = MTS_F32: A and B are fp32 (lda / ldb / batch strides multiples of 4) and the contraction runs on
   * tcgen05 kind::tf32 (operands should already be TF32-representable: mts_round_tf32, round_tf32 below — the tensor
   * core ignores the low 13 mantissa bits); MTS_BF16 (0) = the default bf16 path.
   * round_tf32 != 0 with an fp32 D (not RESID_ADD): store D rounded to nearest TF32, ready to be the next GEMM's operand.
   * With fp32 operands the GELU_NEW / SWIGLU / ROPE_QK epilogues write fp32 D and evaluate expf / tanhf at library
 * accuracy (the bf16 path uses the approximate units, whose error sits below the bf16 output rounding). */
  int32_t ab_dtype;
  int32_t round_tf32;
  /* fp32-grade contraction out of TF32 pieces ("3xTF32"): when a_lo and b_lo are given (fp32 operands only; same
   * shapes / strides as a and b), the operands are A = a + a_lo, B = b + b_lo with every piece TF32-representable
   * (mts_split_tf32) and the kernel accumulates a*b + a_lo*b + a*b_lo into one fp32 accumulator — relative error
   * ~2^-21 instead of TF32's 2^-11 at three times the tensor work.  NULL (both) = plain TF32. */
  const void* a_lo;
  const void* b_lo;
  /* RESID_ADD only: D = C + dropout(alpha * A B^T + bias) with keep probability 1 - drop_p and the counter-based mask
   * of mts_dropout (element (row, col) of the [m, n] result has index row*n + col): the residual dropouts that stay
   * live in the reference's train mode for GPT-2 backbones (tasks/forecasting.py:18 flips the HF module;
   * HF:models/gpt2/modeling_gpt2.py:233, :243 resid_dropout).  0 = off. */
  float drop_p;
  uint64_t drop_seed;
  /* Stream-K (optional, batch 1, bf16 or fp32 operands): when the output tiles do not fill the SMs evenly (small m:
   * M = 800 .. 900 rows of the Ventilator / PSM configs leave 24 .. 60 % of the SMs idle or waiting on a last wave), the
   * library deals the (tile, k-block) units out evenly over one CTA per SM instead; CTAs whose range starts inside a tile
   * park their fp32 partial in `sk_workspace` first thing, the CTA holding the tile's first k-blocks adds them in
   * ascending CTA order (deterministic) at the end of its own range and runs the epilogue.  The caller lends the scratch: sk_workspace >= #SMs * 128 * 256 * 4 bytes,
   * sk_flags >= #SMs int32 (zero-initialised once; every flag raised by a launch is lowered again by its single reader,
   * so launches and graph replays on one stream can share them), sk_epoch the non-zero "ready" value.  NULL workspace = never.  mts_set_option("streamk", 0 off (default) | 1 auto | 2 whenever legal): experimental — on B200 the
   * plain whole-tile schedule is faster at every BASELINE shape (profiles/r02_streamk_gemm.md). */
  void* sk_workspace;
  int64_t sk_workspace_bytes;
  int32_t* sk_flags;
  int32_t sk_flags_len;
  int32_t sk_epoch;
} mts_gemm_args;

int mts_gemm(const mts_gemm_args* args, mts_stream_t stream);

/* Pack [gate; up] weights for MTS_EPI_SWIGLU: out[(j/128)*256 + (j%128)] = gate[j],
 * out[(j/128)*256 + 128 + (j%128)] = up[j]; rows beyond I in the last block are zero.
 *   gate, up: bf16 [I, K];  out: bf16 [2*ceil(I/128)*128, K] */
int mts_pack_gate_up(const uint16_t* gate, const uint16_t* up, uint16_t* out, int I, int K,
                     mts_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Element-wise / row kernels of the backbone                                                  */
/* ------------------------------------------------------------------------------------------ */

/* fp32 -> bf16 cast (weights when the trainer updates them; activations entering a GEMM). */
int mts_cast_f32_bf16(const float* in, uint16_t* out, int64_t n, mts_stream_t stream);
int mts_cast_bf16_f32(const uint16_t* in, float* out, int64_t n, mts_stream_t stream);
/* out[c, r] = in[r, c] with cast; in fp32 [rows, cols] -> out bf16 [cols, rows] */
int mts_transpose_f32_bf16(const float* in, uint16_t* out, int rows, int cols, mts_stream_t stream);
int mts_transpose_bf16(const uint16_t* in, uint16_t* out, int rows, int cols, mts_stream_t stream);

/* K6  RMSNorm (ref: HF:models/llama/modeling_llama.py:53-67): y = w * x * rsqrt(mean(x^2)+eps)
 *   x fp32 [rows, ldx>=D] (the residual stream), w fp32 [D], y bf16 [rows, D] and/or y_f32. */
int mts_rmsnorm(const float* x, int64_t ldx, const float* w, uint16_t* y_bf16, float* y_f32,
                int rows, int D, float eps, mts_stream_t stream);
/* K6  LayerNorm (ref: HF:models/gpt2/modeling_gpt2.py:252-254,628; torch.nn.LayerNorm) */
int mts_layernorm(const float* x, int64_t ldx, const float* w, const float* b, uint16_t* y_bf16,
                  float* y_f32, int rows, int D, float eps, mts_stream_t stream);

/* K8  causal self-attention, eager semantics (ref: HF:models/llama/modeling_llama.py:199-221 with
 *     RoPE :124-168; HF:models/gpt2/modeling_gpt2.py:54-72): softmax(q k^T * scale + causal) v,
 *     no padding mask (the reference never passes one, models/medtsllm.py:350).
 *   qkv  bf16 [Bp*L, 3*H*hd]   columns = [q heads | k heads | v heads]
 *   rope_cos, rope_sin fp32 [L, hd/2] or NULL (GPT-2); rotate-half convention
 *   out  bf16 [Bp*L, H*hd]
 *   lse  fp32 [Bp, H, L] or NULL: log-sum-exp of the scaled scores (saved for backward)
 *   hd in {64, 128}.
 *   Pre-rotated q / k (rope NULL) and L <= 256: runs on tcgen05 with the score tile in tensor memory (TMA loads, Q K^T and
 *   P V as tcgen05.mma, single-pass softmax straight out of TMEM; csrc/attention_tc.cu) unless mts_set_option("attn_tc")
 *   says otherwise; longer sequences and the rotate-while-staging form use the mma.sync kernels. */
int mts_attn_causal(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                    uint16_t* out, float* lse, int Bp, int L, int H, int hd, float scale,
                    mts_stream_t stream);
/* Same attention with dropout on the probabilities after the softmax (train mode of the frozen backbone: GPT-2
 * attn_pdrop, HF:models/gpt2/modeling_gpt2.py:67-68; Llama attention_dropout, HF:models/llama/modeling_llama.py:217).
 * Counter-based mask as mts_dropout: element (b, h, q, k) has index ((b*H + h)*L + q)*L + k.  q / k already rotated,
 * plain row layout, sequences that fit the sequence-resident kernels.  lse is of the UNMASKED scores. */
int mts_attn_causal_dropout(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int L, int H, int hd, float scale,
                            float p, uint64_t seed, mts_stream_t stream);
/* Its backward (rope tables only rotate dq / dk back; NULL for GPT-2): dP = mask * (dO V^T), dV = (mask * P)^T dO. */
int mts_attn_causal_dropout_bwd(const uint16_t* qkv, const float* rope_cos, const float* rope_sin, const uint16_t* out,
                                const uint16_t* dout, const float* lse, float* delta, uint16_t* dqkv, int Bp, int L,
                                int H, int hd, float scale, float p, uint64_t seed, mts_stream_t stream);
/* RoPE applied in place to the q and k sections of qkv (same tables / convention as above).  After it,
 * call mts_attn_causal with NULL tables: short sequences (K and V of one head fit in shared memory)
 * then take the single-staging kernel. */
int mts_rope_qk(uint16_t* qkv, const float* rope_cos, const float* rope_sin, int Bp, int L, int H,
                int hd, mts_stream_t stream);

/* Shared-prefix row layout (in-batch prompt de-duplication).  The reference embeds the same dataset / task
 * prompt in front of every sample (models/medtsllm.py:386-439, :330-339) and, the backbone mask being causal
 * with no padding mask (:350), the states of those Lc leading positions are identical for all samples at every
 * layer.  They are therefore kept ONCE: rows [0, Lc) = the shared prefix (positions 0..Lc-1), rows
 * Lc + b*Ls + t = token t of sample b at position Lc + t (Ls = L - Lc own tokens per sample, Bp samples,
 * Lc + Bp*Ls rows in total).  Every row-wise kernel and GEMM works on that layout unchanged (see
 * mts_gemm_args.rope_prefix); only attention needs to know about it:
 *   qkv bf16 [Lc + Bp*Ls, 3*H*hd] with q / k ALREADY rotated (MTS_EPI_ROPE_QK or mts_rope_qk_shared);
 *   out bf16 [Lc + Bp*Ls, H*hd];  lse fp32 [H*Lc + Bp*H*Ls] (prefix [H, Lc] first, then [Bp, H, Ls]) or NULL.
 * Up to 256 positions the forward runs on tcgen05 with the score tile in tensor memory (csrc/attention_tc.cu, see
 * mts_attn_causal); beyond that all Lc + Ls positions of one head must fit in shared memory (<= ~350 positions at hd 128,
 * ~700 at hd 64) for the mma.sync kernels. */
int mts_attn_causal_shared(const uint16_t* qkv, uint16_t* out, float* lse, int Bp, int Lc, int Ls, int H,
                           int hd, float scale, mts_stream_t stream);
/* Backward for the samples' own tokens only (the prefix has no trainable ancestor when the backbone is frozen):
 * qkv as above (all rows); out_own / dout_own bf16 [Bp*Ls, H*hd], lse_own / delta fp32 [Bp, H, Ls],
 * dqkv_own bf16 [Bp*Ls, 3*H*hd] = the rows from Lc on.  rope tables fp32 [>= Lc+Ls, hd/2] (rotate dq/dk back) or NULL.
 * Lc <= 128, Ls <= 128 (prefix rounded up to 64 + own rows rounded up to 16 <= 256 key columns): one tcgen05 kernel —
 * S = Q K^T and dP = dO V^T in tensor memory, dS / P to swizzled shared memory, dQ = dS K, dV = P^T dO, dK = dS^T Q with the
 * operands consumed in place (MN-major descriptors) — unless mts_set_option("attn_tc", 0); `delta` is then only written
 * when the full backward below asks for it. */
int mts_attn_causal_shared_bwd(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                               const uint16_t* out_own, const uint16_t* dout_own, const float* lse_own,
                               float* delta, uint16_t* dqkv_own, int Bp, int Lc, int Ls, int H, int hd,
                               float scale, mts_stream_t stream);
/* Backward for ALL rows of the shared-prefix layout (something trainable sits inside the backbone: LoRA): the
 * prefix rows receive gradient through the K / V they contribute to every sample's attention.  out, dout bf16
 * [Lc + Bp*Ls, H*hd]; lse, delta fp32 [H*Lc + Bp*H*Ls] laid out as mts_attn_causal_shared writes lse; dqkv bf16
 * [Lc + Bp*Ls, 3*H*hd].  dK / dV of the prefix keys sum over every query row of the batch (one CTA per head and
 * 128 keys sweeps them; no atomics). */
int mts_attn_causal_shared_bwd_full(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                                    const uint16_t* out, const uint16_t* dout, const float* lse, float* delta,
                                    uint16_t* dqkv, int Bp, int Lc, int Ls, int H, int hd, float scale,
                                    mts_stream_t stream);
int mts_rope_qk_shared(uint16_t* qkv, const float* rope_cos, const float* rope_sin, int Bp, int Lc, int Ls,
                       int H, int hd, mts_stream_t stream);

/* row softmax with scale: p = softmax(scale * s) over the last axis.
 *   s fp32 [rows, n], p bf16 [rows, n]   (ref: models/medtsllm.py:587, reprogramming scores) */
int mts_softmax_rows(const float* s, uint16_t* p, int64_t rows, int n, float scale,
                     mts_stream_t stream);

/* K5  prompt gather + left padding + (GPT-2) position embedding, building the backbone input
 *     (ref: models/medtsllm.py:299-311 encode_text/pad_sequence, :331-337, :349 cat;
 *      HF:models/gpt2/modeling_gpt2.py:584-585 wpe add).
 *   ids  int32 [B, Lp]  (already left-padded with the pad id by the host); a NEGATIVE id marks a position that is
 *        not a token — a time-series example part of the prompt (`encode_part`'s tensor branch, models/medtsllm.py:
 *        313-319): its row is written as zero (+wpe) and filled by the reprogramming out-projection afterwards
 *   emb  fp32 [V, D] input-embedding table
 *   wpe  fp32 [>=L, D] or NULL
 *   x    fp32 [B*rep, L, D] out: rows [0,Lp) = emb[ids] (+wpe), rows [Lp,L) = 0 (+wpe);
 *        sample b is written to rows b*rep .. b*rep+rep-1 (repeat_interleave, :343-344). */
int mts_prompt_gather(const int32_t* ids, const float* emb, const float* wpe, float* x, int B,
                      int rep, int Lp, int L, int D, mts_stream_t stream);
/* Same, shared-prefix row layout: the first Lc (<= Lp) prompt tokens are identical in every row of `ids`
 * (host-checked) and are written once to x rows [0, Lc); sample b, replica r owns rows
 * Lc + (b*rep + r)*(L-Lc) + [0, L-Lc).  x fp32 [Lc + B*rep*(L-Lc), D]. */
int mts_prompt_gather_shared(const int32_t* ids, const float* emb, const float* wpe, float* x, int B,
                             int rep, int Lp, int L, int Lc, int D, mts_stream_t stream);

/* Numbers of the "input statistics" prompt (ref: models/medtsllm.py:441-495 build_input_stats_prompt, :530-538
 * calcute_lags), for features [f0, f0 + C_sel) of x fp32 [B, T, C]:
 *   stats fp32 [B, C_sel, 4] = (min, max, lower median as torch.median, trend: 1 if sum(diff(x)) > 0 else 0)
 *   corr  fp64 [B, C_sel, T] workspace: circular autocorrelation sum_t x[t] x[(t+k) mod T] (= irfft(rfft conj rfft) for
 *         even T; the caller keeps the reference's FFT route for odd T, where irfft returns T-1 points)
 *   lags  int32 [B, n_lags]: indices of the n_lags largest values of mean_c corr[b, c, :], descending; the pairs
 *         corr[k] == corr[T-k], which the reference's FFT orders by rounding noise, resolve to the smaller index
 * Two launches; the host reads stats and lags back in one copy instead of the reference's five `.tolist()` syncs. */
int mts_input_stats(const float* x, float* stats, double* corr, int32_t* lags, int B, int T, int C, int f0, int C_sel,
                    int n_lags, mts_stream_t stream);

/* y = silu(g) * u on bf16, g/u being the two halves of a [rows, 2*I] (ld = ldgu) buffer laid out
 * [g | u];  (training path keeps g,u for the backward; ref HF llama :182-184) */
int mts_swiglu(const uint16_t* gu, int64_t ldgu, uint16_t* y, int64_t rows, int I,
               mts_stream_t stream);

/* Dropout for the training path (ref: models/layers/embed.py:183,197 PatchEmbedding.dropout; models/medtsllm.py:587
 * reprogramming attention dropout): y[i] = keep(seed, i) ? x[i] / (1 - p) : 0, counter-based so that the backward
 * re-creates the mask from the same seed (call it on the gradient).  x == y allowed.  dtype: mts_dtype. */
int mts_dropout(const void* x, void* y, int dtype, int64_t n, float p, uint64_t seed, mts_stream_t stream);

/* eval-only output activations, in place on fp32 (ref: models/medtsllm.py:251-259):
 *   sigmoid (binary semantic segmentation / boundary prediction), softmax over n classes. */
int mts_sigmoid(float* y, int64_t n, mts_stream_t stream);
int mts_softmax_lastdim(float* y, int64_t rows, int n, mts_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Evaluation parity modes (ABI 2): fp32 activations, contractions on tcgen05 kind::tf32         */
/* ------------------------------------------------------------------------------------------ */
/* The reference evaluates with fp32 weights and TF32 matmuls (tasks/base.py:19-22; `setup.dtype` "mixed" and
 * "float32" alike).  "tf32" mode follows that regime: every GEMM is mts_gemm with ab_dtype = MTS_F32 on operands
 * rounded to nearest TF32; "fp32" mode runs the same kernels with the 3xTF32 split (a_lo / b_lo), i.e. fp32-grade
 * contractions.  Norms, softmax, RoPE, SwiGLU / GELU and attention are computed in fp32 in both. */
/* y[i] = nearest TF32-representable value of x[i] (y == x allowed). */
int mts_round_tf32(const float* x, float* y, int64_t n, mts_stream_t stream);
/* x = hi + lo with both pieces TF32-representable (hi = tf32(x), lo = tf32(x - hi)): the operands of a 3xTF32 GEMM. */
int mts_split_tf32(const float* x, float* hi, float* lo, int64_t n, mts_stream_t stream);
/* fp32 row softmax with scale: p = softmax(scale * s) over the last axis; s, p fp32 [rows, n]
 * (ref: models/medtsllm.py:587, reprogramming scores). */
int mts_softmax_rows_f32(const float* s, float* p, int64_t rows, int n, float scale, mts_stream_t stream);
/* fp32 causal self-attention, eager semantics (ref: HF:models/llama/modeling_llama.py:199-221,
 * HF:models/gpt2/modeling_gpt2.py:54-72; no padding mask, models/medtsllm.py:350), plain FMA arithmetic.
 *   qkv fp32 [Lc + Bp*Ls, 3*H*hd] (q / k already rotated for Llama), out fp32 [Lc + Bp*Ls, H*hd]
 *   Lc = 0: plain layout, Ls = L.  Lc > 0: shared-prefix row layout (see mts_attn_causal_shared).
 *   round_out != 0: store the result rounded to TF32 (operand of the out-projection in "tf32" mode).  hd in {64, 128}. */
int mts_attn_causal_f32(const float* qkv, float* out, int Bp, int Lc, int Ls, int H, int hd, float scale,
                        int round_out, mts_stream_t stream);
/* The same attention with both contractions on the tensor cores in TF32 (mma.sync m16n8k8; q / k / v / P rounded to
 * nearest TF32, softmax and accumulation fp32): the attention of the "tf32" mode — in the reference's evaluation regime
 * Q K^T and P V are TF32 matmuls as well (tasks/base.py:19-22).  Same arguments and layouts as mts_attn_causal_f32. */
int mts_attn_causal_tf32(const float* qkv, float* out, int Bp, int Lc, int Ls, int H, int hd, float scale,
                         int round_out, mts_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Training path (adapter gradients; dgrad through the frozen backbone)                        */
/* ------------------------------------------------------------------------------------------ */
/* The reference trains patch-embedding, mapping, reprogramming, down-sample and head parameters
 * (all 13-15 adapter tensors; the backbone is frozen, models/medtsllm.py:231-233), so loss.backward()
 * (tasks/forecasting.py:26) flows through every frozen block.  GEMM-shaped pieces of the backward
 * are mts_gemm on transposed operands; the kernels below are the rest.                           */

/* dx (+)= d(RMSNorm)/dx^T (w*dy) and the LayerNorm analogue; x fp32 [rows, ldx], dy bf16 [rows, D],
 * dx fp32 [rows, D]; accumulate != 0 adds into dx (residual-stream gradient).  dx_bf16 (optional, [rows, D]):
 * bf16 copy of the updated dx = the A operand of the next dgrad GEMM (saves a separate cast pass). */
int mts_rmsnorm_bwd(const float* x, int64_t ldx, const float* w, const uint16_t* dy, float* dx,
                    uint16_t* dx_bf16, int rows, int D, float eps, int accumulate, mts_stream_t stream);
int mts_layernorm_bwd(const float* x, int64_t ldx, const float* w, const uint16_t* dy, float* dx,
                      uint16_t* dx_bf16, int rows, int D, float eps, int accumulate, mts_stream_t stream);

/* Causal attention backward (ref: autograd of HF eager attention, HF:models/llama/modeling_llama.py:
 * 199-221).  qkv/out/lse as produced by mts_attn_causal; dout bf16 [Bp*L, H*hd]; delta fp32 [Bp,H,L]
 * workspace; dqkv bf16 [Bp*L, 3*H*hd] receives d(q|k|v) w.r.t. the (un-rotated) projection outputs.
 * pre_roped != 0: qkv already holds rotated q/k (mts_rope_qk); tables are then only used to rotate
 * dq/dk back. */
int mts_attn_causal_bwd(const uint16_t* qkv, const float* rope_cos, const float* rope_sin,
                        const uint16_t* out, const uint16_t* dout, const float* lse, float* delta,
                        uint16_t* dqkv, int Bp, int L, int H, int hd, float scale, int pre_roped,
                        mts_stream_t stream);

/* SwiGLU on saved pre-activations.  Activation column j reads gate column (j/blk)*2*blk + j%blk and
 * the up column blk further: blk = 128 for the packed layout (mts_pack_gate_up), blk = I for [g|u].
 *   gu, dgu bf16 [rows, ld];  y, dact bf16 [rows, I] */
int mts_swiglu_blk(const uint16_t* gu, int64_t ld, uint16_t* y, int64_t rows, int I, int blk,
                   mts_stream_t stream);
int mts_swiglu_bwd(const uint16_t* gu, int64_t ld, const uint16_t* dact, uint16_t* dgu, int64_t rows,
                   int I, int blk, mts_stream_t stream);
/* dact == NULL: out = gelu_new(pre); else out = dact * gelu_new'(pre)   (bf16, n even) */
int mts_gelu_new(const uint16_t* pre, const uint16_t* dact, uint16_t* out, int64_t n,
                 mts_stream_t stream);
/* ds = scale * p * (dp - rowsum(dp*p));  p bf16, dp fp32, ds bf16, all [rows, n] */
int mts_softmax_bwd_rows(const uint16_t* p, const float* dp, uint16_t* ds, int64_t rows, int n,
                         float scale, mts_stream_t stream);
/* out[c] = sum_r x[r*ld + c]  (bias gradients); dtype: mts_dtype of x */
int mts_colsum(const void* x, int dtype, int64_t ld, float* out, int rows, int cols,
               mts_stream_t stream);
/* out[r] = sum_c x[r*ld + c]  (fp32; mapping-layer bias gradient) */
int mts_rowsum_f32(const float* x, int64_t ld, float* out, int rows, int cols, mts_stream_t stream);
/* out[c*ld_out + b*rows + r] = bf16(in[b*in_batch_stride + r*ld_in + c]); dtype: mts_dtype of in */
int mts_transpose_strided(const void* in, int dtype, int64_t ld_in, int64_t in_batch_stride,
                          uint16_t* out, int64_t ld_out, int batch, int rows, int cols,
                          mts_stream_t stream);
/* out[b*out_batch_stride + r*ld_out + c] = bf16(in[b*in_batch_stride + r*ld_in + c])
 * (out_batch_stride 0 = rows*ld_out, i.e. densely stacked batches) */
int mts_cast_rows_f32_bf16(const float* in, int64_t ld_in, int64_t in_batch_stride, uint16_t* out,
                           int64_t ld_out, int64_t out_batch_stride, int batch, int rows, int cols,
                           mts_stream_t stream);
/* Covariate merges over the feature axis (ref: models/medtsllm.py:284-295 `add` / `weighted-average`,
 * :369-377 `independent` / `merge-end`).
 *   group_reduce:  out[b*out_bs + r] = sum_c w[c]*in[(b*C+c)*R + r] + bias[0]   (w NULL: mean; bias may be NULL)
 *   group_reduce_bwd: din = w[c]*dout (or dout/C); optionally dw[c], dbias[0] (needs `in`)
 *   merge_end:     y[b,p,o] = sum_{o2,c} W[o, o2*C+c] * h[b,c,p,o2] + bias[o]    (feature_weighting Linear) */
int mts_group_reduce(const float* in, const float* w, const float* bias, float* out, int64_t out_bs, int B,
                     int C, int64_t R, int accumulate /* out += ... */, mts_stream_t stream);
int mts_group_reduce_bwd(const float* dout, int64_t dout_bs, const float* w, const float* in, float* din,
                         float* dw, float* dbias, int B, int C, int64_t R, mts_stream_t stream);
int mts_merge_end(const float* h, const float* W, const float* bias, float* y, int B, int C, int P, int O,
                  mts_stream_t stream);
int mts_merge_end_bwd(const float* dy, const float* h, const float* W, float* dh, float* dW, float* dbias, int B,
                      int C, int P, int O, mts_stream_t stream);
/* out = dy * stdev[b,c]  (backward of RevIN denorm; statistics are detached, RevIN.py:42-43) */
int mts_revin_denorm_bwd(const float* dy, const float* stdev, float* out, int B, int T, int C,
                         mts_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MTS_B200_H_ */
