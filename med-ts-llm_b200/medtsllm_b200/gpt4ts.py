"""`GPT4TS` — drop-in for the reference class of the same name (models/gpt4ts.py), the model of
BASELINE.json configs[0] (ETTh1 forecasting on the first `gpt_layers` blocks of a frozen GPT-2), on the same
GPT-2 kernel stack as MedTsLLM.

Same constructor `(config, dataset)`, same `forward(inputs: dict) -> Tensor`, same config keys
(`models.gpt4ts.{d_ff,d_model,gpt_layers,train_mlp,patching}`), same parameter names for the model's own
tensors and for the GPT-2 tensors the reference trains (every LayerNorm and the position table,
models/gpt4ts.py:47-53; `train_mlp = true` is not implemented).  Inference and training
(`loss.backward()` through one autograd.Function with a manual backward, like train.py).  Tasks: the four the reference's
Trainers can run it on (forecasting, anomaly_detection, semantic_segmentation, segmentation); like the
reference, `reconstruction` is listed in `supported_tasks` but rejected by `forward` (models/gpt4ts.py:103-104).

Kernels: mts_gpt4ts_embed (normalisation + time-axis TokenEmbedding conv + sinusoid table, csrc/frontend.cu),
the tcgen05 GEMM for the time-axis Linear (batched, bias along M, accumulated onto the wpe rows) and the output
Linear, KernelBackbone for the GPT-2 blocks, mts_revin_denorm for the de-normalisation.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import BIAS_M, BIAS_N, EPI_RESID_ADD, MtsError
from .backbone import KernelBackbone
from .graph import GraphReplay
from .model import _get


class _TokenEmbedding(nn.Module):
    """models/layers/embed.py:29-42 (parameter holder)."""

    def __init__(self, c_in, d_model):
        super().__init__()
        self.tokenConv = nn.Conv1d(c_in, d_model, kernel_size=3, padding=1, padding_mode="circular", bias=False)
        nn.init.kaiming_normal_(self.tokenConv.weight, mode="fan_in", nonlinearity="leaky_relu")


class _PositionalEmbedding(nn.Module):
    """models/layers/embed.py:8-27."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        position = torch.arange(0, max_len).float().unsqueeze(1)
        div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        pe = torch.zeros(max_len, d_model).float()
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class _TimeFeatureEmbedding(nn.Module):
    """models/layers/embed.py:96-106, freq "h" (only used when the dataset provides x_mark)."""

    def __init__(self, d_model):
        super().__init__()
        self.embed = nn.Linear(4, d_model, bias=False)


class _DataEmbedding(nn.Module):
    """models/layers/embed.py:109-120 (parameter holder)."""

    def __init__(self, c_in, d_model):
        super().__init__()
        self.value_embedding = _TokenEmbedding(c_in, d_model)
        self.position_embedding = _PositionalEmbedding(d_model)
        self.temporal_embedding = _TimeFeatureEmbedding(d_model)


class GPT4TS(nn.Module):

    supported_tasks = ["forecasting", "imputation", "reconstruction", "anomaly_detection", "classification",
                       "semantic_segmentation", "segmentation"]
    supported_modes = ["multivariate", "univariate"]

    def __init__(self, config, dataset, backbone: KernelBackbone | None = None):
        """`backbone`: optional injection point (tests / benches); by default the GPT-2 is loaded exactly as the
        reference does, `GPT2Model.from_pretrained("gpt2")` truncated to `gpt_layers` (models/gpt4ts.py:44-45)."""
        super().__init__()
        self.config = config
        mc = self.model_config = _get(_get(config, "models"), "gpt4ts")
        self.task = _get(config, "task")
        self.d_ff, self.d_model = _get(mc, "d_ff"), _get(mc, "d_model")
        self.gpt_layers, self.train_mlp = _get(mc, "gpt_layers"), _get(mc, "train_mlp")
        self.enc_in = self.c_out = dataset.n_features
        self.num_class = dataset.n_classes if self.task in ("classification", "semantic_segmentation") else 0
        self.seq_len = _get(config, "history_len")
        if self.task == "forecasting":
            self.pred_len = _get(config, "pred_len")
        else:
            assert _get(config, "pred_len") == self.seq_len
            self.pred_len = 0
        patching = _get(mc, "patching")
        self.patch_size, self.stride = _get(patching, "patch_len"), _get(patching, "stride")
        self.enc_embedding = _DataEmbedding(self.enc_in * self.patch_size, self.d_model)
        if self.task == "forecasting":
            self.predict_linear_pre = nn.Linear(self.seq_len, self.pred_len + self.seq_len)
            self.predict_linear = nn.Linear(self.patch_size, self.enc_in)
            self.ln = nn.LayerNorm(self.d_ff)
            self.out_layer = nn.Linear(self.d_ff, self.c_out)
        elif self.task in ("anomaly_detection", "reconstruction"):
            self.ln_proj = nn.LayerNorm(self.d_ff)
            self.out_layer = nn.Linear(self.d_ff, self.c_out, bias=True)
        elif self.task == "semantic_segmentation":
            self.ln_proj = nn.LayerNorm(self.d_ff)
            self.out_layer = nn.Linear(self.d_ff, self.num_class if self.num_class > 2 else 1, bias=True)
        elif self.task == "segmentation":
            self.seg_mode = _get(_get(_get(config, "tasks"), "segmentation"), "mode")
            self.ln_proj = nn.LayerNorm(self.d_ff)
            self.out_layer = nn.Linear(self.d_ff, 1, bias=True)
        else:
            raise NotImplementedError(f"GPT4TS task {self.task!r} (imputation / classification) is outside the "
                                      "tasks the reference's Trainers run")
        if self.enc_in * self.patch_size != self.enc_in and self.task != "anomaly_detection":
            # the reference's conv is declared with C*patch_len in-channels but fed C channels
            # (models/gpt4ts.py:41, :136): it only runs with patch_len = 1, as every shipped config sets
            raise ValueError("GPT4TS needs patching.patch_len = 1 (the reference's conv shape mismatch otherwise)")
        self._backbone = backbone
        object.__setattr__(self, "_hf_model", None)
        if backbone is not None and backbone.device.type == "cuda":
            self._mirror_gpt2_parameters()
        if backbone is None:
            from transformers.models.gpt2.modeling_gpt2 import GPT2Model
            hf = GPT2Model.from_pretrained("gpt2", output_attentions=True, output_hidden_states=True)
            hf.h = hf.h[: self.gpt_layers]
            hf.config.n_layer = len(hf.h)
            object.__setattr__(self, "_hf_model", hf)
        hf_cfg = getattr(getattr(self, "_hf_model", None), "config", None)
        self._hf_pdrop = {k: float(getattr(hf_cfg, k + "_pdrop", 0.0) or 0.0) if hf_cfg is not None else 0.0
                          for k in ("embd", "attn", "resid")}
        self.device = None
        self._w_cache: dict[str, tuple] = {}
        self.use_cuda_graph = os.environ.get("MTS_CUDA_GRAPH", "1") != "0"      # see graph.GraphReplay
        self._graph = GraphReplay()

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        dev = self.out_layer.weight.device
        if dev.type == "cuda":
            if self._backbone is None:
                self._backbone = KernelBackbone.from_hf(self._hf_model, dev)
                object.__setattr__(self, "_hf_model", None)
                self._mirror_gpt2_parameters()
            elif self._backbone.device != dev:
                raise MtsError("the kernel backbone lives on another device")
            self.device = dev
        return self

    def _bf16_weight(self, name, p):
        key = (p._version, p.data_ptr())
        hit = self._w_cache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        w = ops.cast_rows(p.detach().contiguous(), rows=p.shape[0], cols=p.shape[1])
        self._w_cache[name] = (key, w)
        return w

    def forward(self, inputs):
        x = inputs["x_enc"]
        if inputs.get("x_mark_enc", None) is not None:
            raise NotImplementedError("x_mark_enc (time-feature embedding) is not provided by the reference's datasets")
        if self.task not in ("forecasting", "anomaly_detection", "semantic_segmentation", "segmentation"):
            raise ValueError("Task name is not valid")                      # models/gpt4ts.py:103-104
        if not x.is_cuda or self._backbone is None:
            raise MtsError("medtsllm_b200.GPT4TS runs on a CUDA device only (no CPU fallback); call .to('cuda')")
        if x.dtype != torch.float32:
            raise MtsError(f"x_enc must be fp32, got {x.dtype}")
        x = x.contiguous()
        B, T, C = x.shape
        assert T == self.seq_len and C == self.enc_in
        bb = self._backbone
        if self.d_model > bb.spec.hidden or self.d_ff > bb.spec.hidden or C > bb.spec.hidden:
            raise MtsError(f"d_model / d_ff / n_features must not exceed the GPT-2 width {bb.spec.hidden}")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            names, params = zip(*self._trainable())
            return _GPT4TSFn.apply(self, x, names, *params)                 # autograd path (training-mode forwards)
        if not self.use_cuda_graph or (self.training and self._dropout_state() is not None):
            return self._forward_impl(x)            # (dropout draws fresh seeds every call)
        key = (tuple(x.shape), x.device.index, self.training, tuple((p._version, p.data_ptr()) for p in self.parameters()))
        return self._graph.run(key, x, self._forward_impl)

    def _dropout_state(self):
        """Train-mode dropouts of the reference's GPT4TS: `DataEmbedding.dropout(training.dropout)` on the embedded window
        (models/layers/embed.py:113, :120-131) and GPT-2's own embd / attn / resid dropouts, live because `model.train()`
        flips the HF module (the "gpt2" checkpoint carries 0.1).  Returns None in evaluation / when everything is zero,
        else {"p_embed", "seed_embed", "bb": backbone dropout dict or None} with fresh seeds from torch's generator.
        `self.backbone_dropout = {"embd", "attn", "resid"}` overrides the HF config (injected backbones carry none)."""
        if not self.training:
            return None
        p_embed = float(_get(_get(self.config, "training"), "dropout", 0.0) or 0.0)
        probs = getattr(self, "backbone_dropout", None) or self._hf_pdrop
        bb_on = max(probs.values()) > 0
        if p_embed <= 0 and not bb_on:
            return None
        n = 2 + 3 * len(self._backbone.layers)
        seeds = [int(v) for v in torch.randint(0, 2 ** 62, (n,))]
        return {"p_embed": p_embed, "seed_embed": seeds[0],
                "bb": {**probs, "seeds": seeds[1:]} if bb_on else None}

    # ------------------------------------------------------------------------------------------ parameters
    def _mirror_gpt2_parameters(self):
        """The reference trains GPT-2's LayerNorm and position parameters (models/gpt4ts.py:47-53: every `ln*` and `wpe`
        tensor has requires_grad = True).  They live in the kernel backbone; this exposes them as nn.Parameters UNDER THE
        REFERENCE'S NAMES (gpt2.wpe.weight, gpt2.h.<i>.ln_1.weight, ..., gpt2.ln_f.bias) sharing the backbone's storage,
        so an optimizer step on them is seen by the kernels directly."""
        if self.train_mlp:
            raise NotImplementedError("models.gpt4ts.train_mlp = true (GPT-2 MLP weight gradients) is not implemented")
        bb = self._backbone

        def holder(**tensors):
            m = nn.Module()
            for k, t in tensors.items():
                m.register_parameter(k, nn.Parameter(t, requires_grad=True))
            return m

        g = nn.Module()
        g.wpe = holder(weight=bb.wpe)
        g.h = nn.ModuleList()
        for lay in bb.layers:
            blk = nn.Module()
            blk.ln_1 = holder(weight=lay["ln1"], bias=lay["ln1b"])
            blk.ln_2 = holder(weight=lay["ln2"], bias=lay["ln2b"])
            g.h.append(blk)
        g.ln_f = holder(weight=bb.final_norm_w, bias=bb.final_norm_b)
        self.gpt2 = g

    def _trainable(self):
        return [(n, p) for n, p in self.named_parameters() if p.requires_grad]

    # ------------------------------------------------------------------------------------------ forward
    def _forward_impl(self, x, stash=None):
        """`stash` (dict): keep what the backward needs (training)."""
        B, T, C = x.shape
        bb = self._backbone
        D, dev = bb.spec.hidden, x.device
        T2 = T + self.pred_len
        drop = self._dropout_state()
        self._last_dropout = drop
        p_emb = drop["p_embed"] if drop is not None else 0.0
        if self.task == "anomaly_detection":
            return self._anomaly_detection(x, stash, drop)
        mean = torch.empty(B, C, device=dev, dtype=torch.float32)
        std = torch.empty(B, C, device=dev, dtype=torch.float32)
        w_conv = self.enc_embedding.value_embedding.tokenConv.weight.detach()
        pe = self.enc_embedding.position_embedding.pe[0, :T]
        X = torch.empty(B * T2, D, device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream().cuda_stream
        enc_nt = None
        if self.task == "forecasting":
            ld_t = ops.ceil8(T)
            enc_t = torch.empty(B, self.d_model, ld_t, device=dev, dtype=torch.bfloat16)
            if stash is not None or p_emb > 0:   # non-transposed copy: B operand of the time-axis Linear's weight gradient
                enc_nt = torch.empty(B, T, self.d_model, device=dev, dtype=torch.bfloat16)
            _lib.call("mts_gpt4ts_embed", x.data_ptr(), w_conv.data_ptr(), pe.data_ptr(), 0, mean.data_ptr(),
                      std.data_ptr(), enc_t.data_ptr(), 0, 0 if enc_nt is None else enc_nt.data_ptr(), B, T, C,
                      self.d_model, D, ld_t, 0, 1e-5, stream)
            if p_emb > 0:
                # DataEmbedding.dropout on [B, T, d_model] (embed.py:131); the transposed operand is rebuilt from it
                ops.dropout(enc_nt, p_emb, drop["seed_embed"], out=enc_nt)
                for b in range(B):
                    ops.transpose_strided(enc_nt, rows=T, cols=self.d_model, in_off=b * T * self.d_model,
                                          out=enc_t[b], ld_out=ld_t)
            # rows of X = wpe[t] (+ 0): HF's GPT2Model adds the position table to inputs_embeds (modeling_gpt2.py:584-585)
            ops.prompt_gather(None, bb.embed, bb.wpe, X, rep=1, Lp=0, L=T2, B=B)
            # Linear along time (models/gpt4ts.py:137): X[b, t', :d_model] += W_pre[t', :] . enc[b, :, :] + b_pre[t']
            w_pre = self._bf16_weight("pre", self.predict_linear_pre.weight)          # [T2, ceil8(T)]
            ops.gemm(w_pre, enc_t, X, m=T2, n=self.d_model, k=T, batch=B, lda=w_pre.shape[1], a_bs=0, ldb=ld_t,
                     b_bs=self.d_model * ld_t, ldd=D, d_bs=T2 * D, bias=self.predict_linear_pre.bias.detach(),
                     bias_axis=BIAS_M, epilogue=EPI_RESID_ADD)
        elif p_emb > 0:
            if self.d_model != D:
                raise NotImplementedError("training.dropout > 0 with d_model != GPT-2 width on the segmentation tasks")
            # dropout(embedding) THEN + wpe: the embedding alone first (zero position table), its dropout, then the rows
            # of the position table with the dropped embedding accumulated onto them
            if getattr(self, "_zero_wpe", None) is None or self._zero_wpe.device != dev:
                self._zero_wpe = torch.zeros_like(bb.wpe)
            E = torch.empty(B * T, D, device=dev, dtype=torch.float32)
            _lib.call("mts_gpt4ts_embed", x.data_ptr(), w_conv.data_ptr(), pe.data_ptr(), self._zero_wpe.data_ptr(),
                      mean.data_ptr(), std.data_ptr(), 0, E.data_ptr(), 0, B, T, C, self.d_model, D, 0, 1, 1e-5, stream)
            ops.dropout(E, p_emb, drop["seed_embed"], out=E)
            ops.prompt_gather(None, bb.embed, bb.wpe, X, rep=1, Lp=0, L=T, B=B)
            ops.group_reduce(E, B, 1, T * D, out=X, accumulate=True)
        else:
            _lib.call("mts_gpt4ts_embed", x.data_ptr(), w_conv.data_ptr(), pe.data_ptr(), bb.wpe.data_ptr(),
                      mean.data_ptr(), std.data_ptr(), 0, X.data_ptr(), 0, B, T, C, self.d_model, D, 0, 1, 1e-5, stream)
        out = self._blocks_and_out_layer(X, B, T2, stash, drop)
        if stash is not None:
            stash.update(x=x, mean=mean, std=std, enc_nt=enc_nt, B=B, T2=T2, drop=drop)
        if self.task == "forecasting":
            ops.revin_denorm(out, mean, std)                                           # :146-147
            return out[:, -self.pred_len:, :].contiguous()
        out = out.squeeze(-1)
        if not self.training:
            if self.task == "semantic_segmentation":
                if self.num_class > 2:
                    out = out.reshape(B, self.seq_len, self.num_class)
                    ops.softmax_lastdim_(out)
                else:
                    ops.sigmoid_(out)
            elif self.seg_mode == "boundary-prediction":
                ops.sigmoid_(out)
        return out

    def _blocks_and_out_layer(self, X, B, T2, stash=None, drop=None):
        """GPT-2 blocks + ln_f, then out_layer on the first d_ff features (models/gpt4ts.py:140-143): fp32 [B, T2, n_out]."""
        bb = self._backbone
        D = bb.spec.hidden
        layers = [] if stash is not None else None
        hid, x_final = bb.forward(X, B, T2, stash=layers, dropout=drop["bb"] if drop is not None else None)   # bf16 [B*T2, D], ln_f applied
        n_out = self.out_layer.weight.shape[0]
        w_out = self._bf16_weight("out", self.out_layer.weight)                        # [n_out, ceil8(d_ff)]
        out = torch.empty(B * T2, n_out, device=X.device, dtype=torch.float32)
        ops.gemm(hid, w_out, out, m=B * T2, n=n_out, k=self.d_ff, lda=D, ldb=w_out.shape[1],
                 bias=self.out_layer.bias.detach(), bias_axis=BIAS_N)
        if stash is not None:
            stash.update(layers=layers, hid=hid, x_final=x_final, rows=B * T2)
        return out.view(B, T2, n_out)

    def _anomaly_detection(self, x, stash=None, drop=None):
        """models/gpt4ts.py:151-177.  The statistics are taken over "segments" of seg_num = 1 time step (:155-159):
        mean = x, the centred series is identically zero, stdev = sqrt(1e-5).  The GPT-2 therefore sees the same
        input for every sample — zeros plus its position table — and the prediction is dec * sqrt(1e-5) + x
        (:172-175).  That one sequence goes through the blocks once; the de-normalisation broadcasts it.  With the
        GPT-2's own dropouts live (training) every sample draws its own masks, so all B sequences go through."""
        B, T, C = x.shape
        bb = self._backbone
        Bb = B if (drop is not None and drop["bb"] is not None) else 1
        X = torch.empty(Bb * T, bb.spec.hidden, device=x.device, dtype=torch.float32)
        ops.prompt_gather(None, bb.embed, bb.wpe, X, rep=1, Lp=0, L=T, B=Bb)           # 0 + wpe[t]
        dec = self._blocks_and_out_layer(X, Bb, T, stash, drop)                        # [Bb, T, C]
        out = dec.expand(B, T, C).contiguous()
        std = torch.full((1, B * T * C), float(torch.tensor(1e-5, dtype=torch.float32).sqrt()), device=x.device)
        ops.revin_denorm(out.view(1, 1, -1), x.reshape(1, -1), std)                    # out * sqrt(1e-5) + x
        if stash is not None:
            stash.update(x=x, B=B, T2=T, drop=drop, Bb=Bb)
        return out

    # ------------------------------------------------------------------------------------------ backward
    def _backward_impl(self, st, dout):
        """Manual backward of `_forward_impl` (the reference gets it from autograd): gradients of the model's own
        tensors and of GPT-2's LayerNorm / position parameters, keyed by parameter name.  Every contraction is the
        tcgen05 NT GEMM on transposed operands, as in train.py."""
        bb = self._backbone
        dev = dout.device
        D, dm, dff = bb.spec.hidden, self.d_model, self.d_ff
        B, T, C = st["x"].shape
        T2 = st["T2"]
        f32 = lambda *s_: torch.empty(*s_, device=dev, dtype=torch.float32)          # noqa: E731
        anomaly = self.task == "anomaly_detection"
        n_out = self.out_layer.weight.shape[0]
        grads = {}
        # ---- output side: de-normalisation / slicing / squeeze
        if self.task == "forecasting":
            full = torch.zeros(B, T2, C, device=dev, dtype=torch.float32)
            full[:, -self.pred_len:, :] = dout                                       # only the last pred_len steps are returned
            dy = ops.revin_denorm_bwd(full, st["std"]).view(B * T2, n_out)            # d dec = dout * std
        elif anomaly and st.get("Bb", 1) > 1:
            # every sample went through the blocks (live GPT-2 dropouts): d dec[b] = sqrt(1e-5) * dout[b]
            dy = ops.revin_denorm_bwd(dout.reshape(1, 1, -1).contiguous(),
                                      torch.full((1, B * T * C), float(torch.tensor(1e-5).sqrt()), device=dev)).view(B * T, n_out)
        elif anomaly:
            # out[b] = dec * sqrt(1e-5) + x[b] with ONE dec for the whole batch: d dec = sqrt(1e-5) * sum_b dout[b]
            summed = ops.colsum(dout.reshape(B, T * C).contiguous()).view(1, T * C)
            dy = ops.revin_denorm_bwd(summed.view(1, 1, -1), torch.full((1, T * C), float(torch.tensor(1e-5).sqrt()),
                                                                        device=dev)).view(T, n_out)
        else:
            dy = dout.reshape(B * T2, n_out).contiguous()
        M = dy.shape[0]                                                              # rows through the blocks
        Bb = M // T2                                                                 # sequences through the blocks
        # ---- out_layer: dec = hid[:, :d_ff] W_out^T + b
        grads["out_layer.bias"] = ops.colsum(dy)
        Mp = ops.ceil8(M)
        hid_t = ops.transpose_strided(st["hid"], rows=M, cols=dff, ld_in=D)           # [d_ff, Mp]
        dy_t = ops.transpose_strided(dy, rows=M, cols=n_out)                          # [n_out, Mp]
        g_wout = f32(n_out, dff)
        ops.gemm(dy_t, hid_t, g_wout, m=n_out, n=dff, k=M, lda=Mp, ldb=Mp)
        grads["out_layer.weight"] = g_wout
        dy_b = ops.cast_rows(dy, rows=M, cols=n_out)                                  # [M, ceil8(n_out)]
        w_out = self._bf16_weight("out", self.out_layer.weight)                       # [n_out, ceil8(d_ff)]
        w_out_t = ops.transpose_strided(w_out, rows=n_out, cols=dff, ld_in=w_out.shape[1])   # [d_ff, ceil8(n_out)]
        dhid = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
        ops.gemm(dy_b, w_out_t, dhid, m=M, n=dff, k=n_out, lda=dy_b.shape[1], ldb=w_out_t.shape[1], ldd=D)
        # ---- GPT-2 blocks: dgrad + LayerNorm parameter gradients
        ng = {}
        drop = st.get("drop")
        dX, _ = bb.backward(dhid, st["x_final"], st["layers"], Bb, T2, norm_grads=ng,
                            dropout=drop["bb"] if drop is not None else None)            # fp32 [M, D]
        p_emb = drop["p_embed"] if drop is not None else 0.0
        grads["gpt2.ln_f.weight"], grads["gpt2.ln_f.bias"] = ng[("ln_f",)]
        for li in range(len(bb.layers)):
            grads[f"gpt2.h.{li}.ln_1.weight"], grads[f"gpt2.h.{li}.ln_1.bias"] = ng[(li, "ln1")]
            grads[f"gpt2.h.{li}.ln_2.weight"], grads[f"gpt2.h.{li}.ln_2.bias"] = ng[(li, "ln2")]
        # ---- position table: X[b, t] = emb[b, t] + wpe[t]
        g_pos = ops.colsum(dX.view(Bb, T2 * D)).view(T2, D) if Bb > 1 else dX.view(T2, D)
        g_wpe = torch.zeros_like(bb.wpe)
        g_wpe[:T2] = g_pos
        grads["gpt2.wpe.weight"] = g_wpe
        if anomaly:
            return grads
        # ---- embedding side
        x, mean, std = st["x"], st["mean"], st["std"]
        g_conv = f32(dm, C, 3)
        stream = torch.cuda.current_stream().cuda_stream
        if self.task == "forecasting":
            # X[b, t', :dm] = sum_t W_pre[t', t] enc[b, t, :] + b_pre[t']
            dXb = ops.cast_bf16(dX)                                                   # [B*T2, D]
            g_wpre = f32(T2, T)
            for b in range(B):                                                        # sum over the batch, in place
                ops.gemm(dXb, st["enc_nt"], g_wpre, m=T2, n=T, k=dm, lda=D, a_off=b * T2 * D, ldb=dm, b_off=b * T * dm,
                         epilogue=EPI_RESID_ADD if b else 0)
            grads["predict_linear_pre.weight"] = g_wpre
            grads["predict_linear_pre.bias"] = ops.rowsum(g_pos[:, :dm].contiguous())
            w_pre = self._bf16_weight("pre", self.predict_linear_pre.weight)          # [T2, ceil8(T)]
            w_pre_t = ops.transpose_strided(w_pre, rows=T2, cols=T, ld_in=w_pre.shape[1])     # [T, ceil8(T2)]
            if p_emb > 0:
                # d enc in [B, T, d_model] order (transposed store), so that the DataEmbedding dropout mask applies
                denc = f32(B, T, dm)
                for b in range(B):
                    dx_t = ops.transpose_strided(dX, rows=T2, cols=dm, ld_in=D, in_off=b * T2 * D)   # [dm, ceil8(T2)]
                    ops.gemm(dx_t, w_pre_t, denc[b], m=dm, n=T, k=T2, lda=dx_t.shape[1], ldb=w_pre_t.shape[1],
                             d_transposed=True, ldd=dm)
                ops.dropout(denc, p_emb, drop["seed_embed"], out=denc)
                _lib.call("mts_gpt4ts_conv_wgrad", x.data_ptr(), mean.data_ptr(), std.data_ptr(), denc.data_ptr(),
                          g_conv.data_ptr(), B, T, C, dm, T * dm, 1, dm, stream)
            else:
                denc_t = f32(B, dm, T)                                                # d enc, transposed like enc_t
                for b in range(B):
                    dx_t = ops.transpose_strided(dX, rows=T2, cols=dm, ld_in=D, in_off=b * T2 * D)   # [dm, ceil8(T2)]
                    ops.gemm(dx_t, w_pre_t, denc_t[b], m=dm, n=T, k=T2, lda=dx_t.shape[1], ldb=w_pre_t.shape[1])
                _lib.call("mts_gpt4ts_conv_wgrad", x.data_ptr(), mean.data_ptr(), std.data_ptr(), denc_t.data_ptr(),
                          g_conv.data_ptr(), B, T, C, dm, dm * T, T, 1, stream)
        else:
            # the embedding went straight into the residual stream: d enc[b, t, d] = dX[b, t, d] (x its dropout mask)
            dE = ops.dropout(dX, p_emb, drop["seed_embed"]) if p_emb > 0 else dX
            _lib.call("mts_gpt4ts_conv_wgrad", x.data_ptr(), mean.data_ptr(), std.data_ptr(), dE.data_ptr(),
                      g_conv.data_ptr(), B, T, C, dm, T * D, 1, D, stream)
        grads["enc_embedding.value_embedding.tokenConv.weight"] = g_conv
        return grads


class _GPT4TSFn(torch.autograd.Function):
    """`loss.backward()` of the reference's Trainer (tasks/forecasting.py:26) through GPT4TS on the kernel stack."""

    @staticmethod
    def forward(ctx, model, x, names, *params):
        stash = {}
        out = model._forward_impl(x, stash)
        ctx.model, ctx.stash, ctx.names = model, stash, names
        return out

    @staticmethod
    @torch.no_grad()
    def backward(ctx, dout):
        grads = ctx.model._backward_impl(ctx.stash, dout.float().contiguous())
        ctx.stash = None
        # parameters the forward never reads (temporal_embedding, predict_linear, ln / ln_proj: unused by the
        # reference's forward too) get no gradient, exactly as under autograd
        return (None, None, None) + tuple(grads.get(n) for n in ctx.names)
