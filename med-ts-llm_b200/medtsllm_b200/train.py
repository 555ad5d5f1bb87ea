"""Training path: `loss.backward()` of the reference's Trainer (tasks/forecasting.py:26) through the
kernel stack.

The reference trains every adapter around the frozen LLM — patch-embedding conv, mapping layer,
the four reprogramming projections, the down-sample Linear and the flatten head (15 tensors with the
shipped configs) — and three of those sit IN FRONT of the backbone, so the gradient has to flow
through all frozen blocks (dgrad only, no backbone weight gradients).  One autograd.Function wraps
the whole hot path: forward = MedTsLLM._forward_impl with an activation stash, backward = the manual
chain below.  Every contraction is the tcgen05 NT GEMM (mts_gemm) on transposed operands
(dX = dY W -> B = W^T; dW = dY^T X -> A = dY^T, B = X^T, both produced by our transpose kernel);
the rest are the kernels of csrc/backward.cu and the attention backward.

Numerics: bf16 operands / fp32 accumulation like the forward (= the reference's bf16-autocast
regime); the residual-stream gradient is fp32.  `training.dropout` (patch-embedding and reprogramming-attention
dropout) uses a counter-based mask that the backward re-creates from the step's seeds.
"""
from __future__ import annotations

import torch

from . import dp, ops
from ._lib import EPI_RESID_ADD, MtsError, launch_count, note_replay


def forward_train(model, inputs):
    params = model.adapter_params()
    return _HotPathFn.apply(model, inputs, *params)


class TrainGraph:
    """CUDA-graph replay of the training step's device work: one graph for the forward (with its activation
    stash), one for the backward chain.  The launch-bound configurations (GPT-2 backbones, LoRA: 600-1450 short
    kernels per step) are otherwise limited by the host issuing the launches.

    Key = input shape, prompt-id table, row layout and the ADDRESS of every trainable tensor — not its version:
    the optimizer updates the fp32 masters in place and the captured forward re-casts them to bf16 every replay
    (`model._force_recast` makes the capture include every cast / prototype GEMM even if a cache would hit).
    The gradients come back in static buffers and are handed to autograd as copies.  With data parallelism the
    all-reduces run between the two backward graphs (NCCL stays outside the capture)."""

    def __init__(self):
        self.entry = None
        self.seen = None
        self.generation = 0

    @staticmethod
    def key_of(model, x_enc, ids):
        return (tuple(x_enc.shape), x_enc.device.index, id(ids), model.share_prompt_prefix,
                tuple(p.data_ptr() for p in model.adapter_params()))

    def forward(self, model, inputs, x_enc, ids):
        """Returns (out, True) when the step runs on the graph, (None, False) when the caller must run eagerly."""
        key = self.key_of(model, x_enc, ids)
        e = self.entry
        if e is not None and e["key"] == key:
            e["x"].copy_(x_enc)
            e["fwd"].replay()
            note_replay(e["fwd_launches"])
            self.generation += 1
            return e["out"].clone(), True
        if self.seen is None or self.seen[0] != key:
            self.seen = (key, ids)
            return None, False
        self.entry = None
        static_x = x_enc.clone()
        stash = {}
        graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        model._set_force_recast(True)
        try:
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = model._forward_impl({**inputs, "x_enc": static_x}, stash, ids)
        finally:
            model._set_force_recast(False)
        self.entry = {"key": key, "fwd": graph, "x": static_x, "out": out, "stash": stash, "ids": ids,
                      "fwd_launches": launch_count() - n0, "bwd": None}
        graph.replay()
        self.generation += 1
        return out.clone(), True

    def backward(self, model, dout, generation):
        e = self.entry
        if e is None or generation != self.generation:
            raise MtsError("backward() of a training step whose activations were overwritten by a later forward on "
                           "the captured graph (set model.use_train_graph = False for this usage pattern)")
        # two graphs: everything up to the exchanged gradients, then the mapping-layer gradient from the (averaged)
        # dSource.  The collectives run between them, in the same order as on the kernel-by-kernel path, so ranks
        # that disagree on graph vs eager (a per-rank decision) still issue matching NCCL sequences.
        if e["bwd"] is None:
            e["dout"] = dout.clone()
            graph = torch.cuda.CUDAGraph()
            n0 = launch_count()
            with torch.cuda.graph(graph, pool=e["fwd"].pool(), capture_error_mode="thread_local"):
                e["ctx"] = _backward_part1(model, e["stash"], e["dout"], use_dp=False)
            e["bwd"], e["bwd_launches"] = graph, launch_count() - n0
            graph.replay()
        else:
            e["dout"].copy_(dout)
            e["bwd"].replay()
            note_replay(e["bwd_launches"])
        ctx = e["ctx"]
        active = dp.is_active()
        if active:
            ctx["early"].launch()
            ctx["dsrc_arena"].launch()
            ctx["late"].launch()
            ctx["lora_arena"].launch()
            ctx["dsrc_arena"].finish()
        if e.get("bwd2") is None:
            graph = torch.cuda.CUDAGraph()
            n0 = launch_count()
            with torch.cuda.graph(graph, pool=e["fwd"].pool(), capture_error_mode="thread_local"):
                e["grads"] = _backward_part2(model, ctx)
            e["bwd2"], e["bwd2_launches"] = graph, launch_count() - n0
            graph.replay()
        else:
            e["bwd2"].replay()
            note_replay(e["bwd2_launches"])
        if active:
            ctx["early"].finish()
            ctx["late"].finish()
            ctx["lora_arena"].finish()
        return [g.clone() if g is not None else None for g in e["grads"]]


def _t(x, **kw):
    """[rows, cols] -> bf16 [cols, ceil8(rows)] (zero padded): a K-major operand with K = rows."""
    rows, cols = x.shape
    return ops.transpose_strided(x, rows=rows, cols=cols, **kw)


class _HotPathFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, model, inputs, *params):
        ctx.model = model
        ctx.graph_generation = None
        if model.train_graph_enabled():
            x_enc = model._check_input(inputs)
            ids = model.prompt_token_ids(inputs)
            out, on_graph = model._train_graph.forward(model, inputs, x_enc, ids)
            if on_graph:
                ctx.graph_generation = model._train_graph.generation
                ctx.stash = None
                return out
            stash = {}
            out = model._forward_impl(inputs, stash, ids)
        else:
            stash = {}
            out = model._forward_impl(inputs, stash)
        ctx.stash = stash
        return out

    @staticmethod
    @torch.no_grad()
    def backward(ctx, dout):
        m = ctx.model
        if dout.dtype != torch.float32:
            dout = dout.float()
        if ctx.graph_generation is not None:
            grads = m._train_graph.backward(m, dout.contiguous(), ctx.graph_generation)
        else:
            grads = _backward_chain(m, ctx.stash, dout, use_dp=True)
            ctx.stash = None
        return (None, None) + tuple(grads)


def _arena_layout(m):
    """(early, late) name/shape lists of the adapter gradients.  `early` = final before the backbone dgrad starts
    (head, down-sample, merge-end weighting); `late` = everything in front of the backbone except the mapping layer,
    whose weight gradient is derived AFTER the exchange from the averaged `dsrc` (see dp.py)."""
    named = dict(m.named_parameters())
    mode = m.covariate_mode
    early = ["output_projection.linear.weight", "output_projection.linear.bias"]
    if m.embedding_downsample_mode == "linear":
        early += ["embedding_downsample_layer.weight", "embedding_downsample_layer.bias"]
    if mode == "merge-end":
        early += ["feature_weighting.weight", "feature_weighting.bias"]
    late = ["patch_embedding.value_embedding.tokenConv.weight"]
    for proj in ("query", "key", "value", "out"):
        late += [f"reprogramming_layer.{proj}_projection.weight", f"reprogramming_layer.{proj}_projection.bias"]
    if mode == "weighted-average":
        late += ["feature_weighting.weight", "feature_weighting.bias"]
    shape = lambda names: [(k, tuple(named[k].shape)) for k in names]      # noqa: E731
    return shape(early), shape(late)


def _backward_chain(m, st, dout, use_dp=True):
    """The manual backward of the hot path (see the module docstring): returns the gradients aligned with
    `m.adapter_params()`.  Kernel-by-kernel path; with data parallelism the exchanges (early arena, dSource, late
    arena, LoRA arena — always in this order, the graph path issues the same sequence) overlap the chain."""
    ctx = _backward_part1(m, st, dout, use_dp=use_dp)
    if use_dp:
        ctx["dsrc_arena"].finish()
    grads = _backward_part2(m, ctx)
    if use_dp:
        ctx["early"].finish()
        ctx["late"].finish()
        ctx["lora_arena"].finish()
    return grads


def _backward_part1(m, st, dout, use_dp):
    """Everything up to the gradients that are exchanged between ranks: fills the early / late arenas and `dsrc`
    (gradient of the prototype matrix `source`), launching their all-reduces as they become final when `use_dp`."""
    bb = m._backbone
    dev = dout.device
    B0, B, N, N0, E, H, D = st["B"], st["Bp"], m.n_patches, st["N0"], m.d_ff, m.n_attention_heads, m.d_llm
    HE, S, Lp, L, C = H * E, m.num_tokens, st["Lp"], st["L"], m.n_features
    Lc = st["Lc"]               # shared-prefix rows: the backward runs on the sequences' own rows only ...
    Ls = L - Lc                 # own rows per sequence
    row0 = Lc if (Lc and m.lora_enabled) else 0    # ... unless LoRA needs the prefix rows too (all rows then)
    own_off = Lp - Lc + row0    # first patch row of sequence 0 in dhid / dR
    mode = m.covariate_mode
    R_main = st["enc"].shape[0] * N0            # reprogrammed rows of the windows, ordered (sample, [feature,] patch)
    ex_meta = st.get("ex_meta") or []           # + the patches of time-series example parts inside the prompts
    R = st["enc_all"].shape[0] if ex_meta else R_main
    dm = m.d_model
    rl = m.reprogramming_layer
    f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)     # noqa: E731
    bf = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)     # noqa: E731
    zbf = lambda *s: torch.zeros(*s, device=dev, dtype=torch.bfloat16)    # noqa: E731
    early_l, late_l = _arena_layout(m)
    early, late = dp.GradArena(early_l, dev), dp.GradArena(late_l, dev)
    dsrc_arena = dp.GradArena([("dsrc", (S, D))], dev)
    gv = lambda k: (early if k in early else late).view(k)                # noqa: E731

    # ---- de-norm / squeeze (models/medtsllm.py:379-382): statistics are detached
    dout = dout.reshape(B0, m.pred_len, m.n_outputs_per_step).contiguous()
    dy = ops.revin_denorm_bwd(dout, st["std"]) if st["denorm"] else dout
    n_out = m.n_outputs
    # ---- merges over the feature axis after the head (models/medtsllm.py:369-377)
    if mode == "independent":
        dy2, _, _ = ops.group_reduce_bwd(dy.view(B0, n_out), B0, C, n_out)            # [B0*C, n_out]
    elif mode == "merge-end":
        dy2, g_fw_w, g_fw_b = ops.merge_end_bwd(dy, st["head"], m.feature_weighting.weight.detach(), B0, C,
                                                m.pred_len, m.n_outputs_per_step)
        gv("feature_weighting.weight").copy_(g_fw_w)
        gv("feature_weighting.bias").copy_(g_fw_b)
    else:
        dy2 = dy.view(B, n_out)

    # ---- flatten head: out = flat W_h^T + b_h
    ops.colsum(dy2, out=gv("output_projection.linear.bias"))
    dy_b = ops.cast_rows(dy2, rows=B, cols=n_out)                          # bf16 [B, ceil8(n_out)]
    wh = m._bf16_weight("wh", m.output_projection.linear.weight)           # [n_out, ceil8(EN)]
    EN = E * N
    ops.gemm(_t(dy2), _t(st["flat"]), gv("output_projection.linear.weight"), m=n_out, n=EN, k=B,
             lda=ops.ceil8(B), ldb=ops.ceil8(B))
    wh_t = ops.transpose_strided(wh, rows=n_out, cols=EN, ld_in=wh.shape[1])   # [EN, ceil8(n_out)]
    dflat = bf(B, EN)                                                       # [B, E, N]
    ops.gemm(dy_b, wh_t, dflat, m=B, n=EN, k=n_out, lda=dy_b.shape[1], ldb=wh_t.shape[1])

    # ---- down-sample Linear on the last N tokens: flat[b, f, n] = hid[b, Lp+n] . W_ds[f] + b_ds[f]
    # dflat is [B][E][N]; dY_ds[(b, n), f] = dflat[b, f, n] is a per-batch transpose
    Rh = B * N                                  # rows entering the down-sample step: (sequence, token)
    dyds = bf(Rh, E)
    for b in range(B):   # B small launches of a tiny kernel (B*N*E elements in total)
        ops.transpose_strided(dflat, rows=E, cols=N, in_off=b * EN, out=dyds[b * N:(b + 1) * N], ld_out=E)
    if m.embedding_downsample_mode == "linear":
        ops.colsum(dyds, out=gv("embedding_downsample_layer.bias"))
        hid_last_t = ops.transpose_strided(st["hid"], batch=B, rows=N, cols=D, ld_in=D, in_bs=Ls * D,
                                           in_off=Lp * D)                  # [D, ceil8(R)]
        ops.gemm(_t(dyds), hid_last_t, gv("embedding_downsample_layer.weight"), m=E, n=D, k=Rh,
                 lda=ops.ceil8(Rh), ldb=hid_last_t.shape[1])
    wds, _ = m._downsample_operands()                                       # [E, D] (trainable or constant)
    wds_t = ops.transpose_strided(wds, rows=E, cols=D, ld_in=wds.shape[1])  # [D, ceil8(E)]
    dhid = zbf(row0 + B * Ls, D)                                            # zero for prompt rows
    ops.gemm(dyds, wds_t, dhid, m=N, n=D, k=E, batch=B, a_bs=N * E, b_bs=0, ldb=wds_t.shape[1],
             d_bs=Ls * D, ldd=D, d_off=own_off * D)

    # ---- DP: the head / down-sample gradients are final -> all-reduce them underneath the backbone dgrad
    if use_dp:
        early.launch()

    # ---- frozen backbone (dgrad only)
    dR, lora_grads = bb.backward(dhid, st["x_final"], st["layers"], B, L,
                                 lora=m.llm if m.lora_enabled else None, Lc=Lc, dropout=st.get("bb_drop"))   # fp32 [B*Ls, D]

    # ---- reprogramming out-projection: rows (sample, [feature,] patch) of O W_o^T + b_o feed X's patch rows
    if mode in ("concat", "univariate", "independent", "merge-end"):
        dxp = bf(R, D)
        ops.cast_rows(dR, batch=B, rows=N, cols=D, ld_in=D, in_bs=Ls * D, in_off=own_off * D, out=dxp, ld_out=D)
        for ex in ex_meta:       # gradient rows of the example patches: prompt position `pos` of sample b
            ops.cast_rows(dR, rows=ex["n"], cols=D, ld_in=D, in_off=(row0 + ex["b"] * Ls + ex["pos"] - Lc) * D,
                          out=dxp, ld_out=D, out_off=ex["r0"] * D)
    elif mode == "interleave":
        dxp = bf(R, D)
        for c in range(C):       # feature c owns rows Lp + n*C + c
            ops.cast_rows(dR, batch=B0, rows=N0, cols=D, ld_in=C * D, in_bs=Ls * D, in_off=(own_off + c) * D,
                          out=dxp, ld_out=D, out_bs=C * N0 * D, out_off=c * N0 * D)
    else:                        # add / weighted-average: broadcast back over the feature axis
        fw = m.feature_weighting if mode == "weighted-average" else None
        dyf, g_w, g_b = ops.group_reduce_bwd(dR, B0, C, N0 * D, w=fw.weight.detach().view(-1) if fw is not None else None,
                                             x=st["Y"] if fw is not None else None, dout_bs=Ls * D, dout_off=own_off * D)
        if fw is not None:
            gv("feature_weighting.weight").copy_(g_w.view(1, C))
            gv("feature_weighting.bias").copy_(g_b.view(1))
        dxp = ops.cast_bf16(dyf.view(R, D))
    ops.colsum(dxp, out=gv("reprogramming_layer.out_projection.bias"))
    Rp = ops.ceil8(R)
    ops.gemm(_t(dxp), _t(st["O"]), gv("reprogramming_layer.out_projection.weight"), m=D, n=HE, k=R, lda=Rp, ldb=Rp)
    wo = m._bf16_weight("wo", rl.out_projection.weight)                    # [D, HE]
    wo_t = ops.transpose_strided(wo, rows=D, cols=HE, ld_in=wo.shape[1])   # [HE, D]
    dO = bf(R, HE)
    ops.gemm(dxp, wo_t, dO, m=R, n=HE, k=D, ldb=wo_t.shape[1])

    # ---- cross-attention core, per head h: O_h = P_h V_h, P_h = softmax(scale Q_h K_h^T)
    P, Pd, Q, K, Vt = st["P"], st["Pd"], st["Q"], st["K"], st["Vt"]
    p_drop, seeds = st["p_drop"], st["seeds"]
    Vm = ops.transpose_strided(Vt, rows=HE, cols=S)                        # [S, HE]
    dP = f32(H, R, S)
    ops.gemm(dO, Vm, dP, m=R, n=S, k=E, batch=H, lda=HE, a_bs=E, ldb=Vm.shape[1], b_bs=E, d_bs=R * S)
    if p_drop > 0:
        ops.dropout(dP, p_drop, seeds[1], out=dP)                          # same mask as the forward's attention dropout
    dS = ops.softmax_bwd_rows(P, dP, st["scale"])                          # bf16 [H, R, S]
    # per-head transposes laid out [S, H*Rp]: head h occupies columns [h*Rp, h*Rp + R)
    P_t, dS_t = zbf(S, H * Rp), zbf(S, H * Rp)
    for h in range(H):
        ops.transpose_strided(Pd, rows=R, cols=S, in_off=h * R * S, out=P_t[:, h * Rp:], ld_out=H * Rp)
        ops.transpose_strided(dS, rows=R, cols=S, in_off=h * R * S, out=dS_t[:, h * Rp:], ld_out=H * Rp)
    dO_t, Q_t = _t(dO), _t(Q)                                              # [HE, Rp]
    dV = bf(S, HE)
    ops.gemm(P_t, dO_t, dV, m=S, n=E, k=R, batch=H, lda=H * Rp, a_bs=Rp, ldb=Rp, b_bs=E * Rp, ldd=HE, d_bs=E)
    dK = bf(S, HE)
    ops.gemm(dS_t, Q_t, dK, m=S, n=E, k=R, batch=H, lda=H * Rp, a_bs=Rp, ldb=Rp, b_bs=E * Rp, ldd=HE, d_bs=E)

    # ---- gradient of the prototypes `source` (through the key / value projections): exchanged between ranks in place
    # of the mapping-layer weight gradient derived from it (part 2)
    wk = m._bf16_weight("wk", rl.key_projection.weight)                    # [HE, D]
    wv = m._bf16_weight("wv", rl.value_projection.weight)
    dsrc = dsrc_arena.view("dsrc")
    ops.gemm(dK, _t(wk), dsrc, m=S, n=D, k=HE, ldb=ops.ceil8(HE))
    ops.gemm(dV, _t(wv), dsrc, m=S, n=D, k=HE, ldb=ops.ceil8(HE), epilogue=EPI_RESID_ADD)
    if use_dp:
        dsrc_arena.launch()

    K_t = _t(K)                                                            # [HE, S]
    dQ = bf(R, HE)
    ops.gemm(dS, K_t, dQ, m=R, n=E, k=S, batch=H, a_bs=R * S, lda=S, ldb=K_t.shape[1], b_bs=E * K_t.shape[1],
             ldd=HE, d_bs=E)

    # ---- query projection + front end
    ops.colsum(dQ, out=gv("reprogramming_layer.query_projection.bias"))
    enc2 = st["enc_all"] if ex_meta else st["enc"].view(R, dm)
    ops.gemm(_t(dQ), _t(enc2), gv("reprogramming_layer.query_projection.weight"), m=HE, n=dm, k=R, lda=Rp, ldb=Rp)
    wq = m._bf16_weight("wq", rl.query_projection.weight)                  # [HE, ceil8(dm)]
    wq_t = ops.transpose_strided(wq, rows=HE, cols=dm, ld_in=wq.shape[1])  # [dm, HE]
    denc = f32(R, dm)
    ops.gemm(dQ, wq_t, denc, m=R, n=dm, k=HE, ldb=wq_t.shape[1])
    denc_main = denc[:R_main]
    if p_drop > 0:
        ops.dropout(denc_main, p_drop, seeds[0], out=denc_main)            # patch-embedding dropout mask
        if ex_meta:
            ops.dropout(denc[R_main:], p_drop, seeds[2], out=denc[R_main:])
    g_conv = gv("patch_embedding.value_embedding.tokenConv.weight")
    ops.revin_patch_embed_bwd(st["x_enc"], st["mean"], st["std"], denc_main.view(st["enc"].shape),
                              m.patch_len, m.stride, m.d_patch, concat=st["concat"], out=g_conv)
    for ex in ex_meta:           # the conv weight also sees the example parts (each with its own RevIN statistics)
        g_conv += ops.revin_patch_embed_bwd(ex["ts"], ex["mean"], ex["std"],
                                            denc[ex["r0"]:ex["r0"] + ex["n"]].view(1, ex["n"], dm).contiguous(),
                                            m.patch_len, m.stride, m.d_patch, concat=st["concat"])

    # ---- key / value projections of the prototypes
    source = st["source"]
    src_t = _t(source)                                                     # [D, S]
    ops.colsum(dK, out=gv("reprogramming_layer.key_projection.bias"))
    ops.colsum(dV, out=gv("reprogramming_layer.value_projection.bias"))
    ops.gemm(_t(dK), src_t, gv("reprogramming_layer.key_projection.weight"), m=HE, n=D, k=S, lda=ops.ceil8(S),
             ldb=src_t.shape[1])
    ops.gemm(_t(dV), src_t, gv("reprogramming_layer.value_projection.weight"), m=HE, n=D, k=S, lda=ops.ceil8(S),
             ldb=src_t.shape[1])
    if use_dp:
        late.launch()
    # LoRA pairs: one flat buffer as well (128 small tensors for Llama-2-7B: a `torch.cat` + per-tensor copy back around
    # the all-reduce would cost more launches than the exchange itself; inside a captured step the copies are graph nodes)
    lora_arena = dp.GradArena([(f"lora{i}", tuple(g.shape)) for i, g in enumerate(lora_grads or [])], dev)
    for i, g in enumerate(lora_grads or []):
        lora_arena.view(f"lora{i}").copy_(g)
    lora_views = [lora_arena.view(f"lora{i}") for i in range(len(lora_grads or []))]
    if use_dp:
        lora_arena.launch()
    return {"early": early, "late": late, "dsrc_arena": dsrc_arena, "lora_grads": lora_views, "lora_arena": lora_arena}


def _backward_part2(m, ctx):
    """Mapping layer (models/medtsllm.py:281) from the — under DP already averaged — gradient of `source`:
    d W_map = dSource E, d b_map = rowsum(dSource).  Returns all gradients in `m.adapter_params()` order."""
    bb = m._backbone
    dsrc = ctx["dsrc_arena"].view("dsrc")
    S, V = m.num_tokens, m.vocab_size
    dsrc_b = ops.cast_bf16(dsrc)
    g_bmap = ops.rowsum(dsrc)                                              # d b_map[s] = sum_d dSource[s, d]
    g_wmap = torch.empty(S, V, device=dsrc.device, dtype=torch.float32)
    ops.gemm(dsrc_b, bb.embed_bf16(), g_wmap, m=S, n=V, k=m.d_llm)
    early, late = ctx["early"], ctx["late"]
    out = []
    for k in m.param_order():
        if k == "mapping_layer.weight":
            out.append(g_wmap)
        elif k == "mapping_layer.bias":
            out.append(g_bmap)
        else:
            out.append((early if k in early else late).view(k))
    return out + list(ctx["lora_grads"])
