"""`GPT4TS` — drop-in for the reference class of the same name (models/gpt4ts.py), the model of
BASELINE.json configs[0] (ETTh1 forecasting on the first `gpt_layers` blocks of a frozen GPT-2), on the same
GPT-2 kernel stack as MedTsLLM.

Same constructor `(config, dataset)`, same `forward(inputs: dict) -> Tensor`, same config keys
(`models.gpt4ts.{d_ff,d_model,gpt_layers,train_mlp,patching}`), same parameter names for the model's own
tensors.  Inference only in this round: the reference also trains GPT-2's LayerNorm / position parameters
(models/gpt4ts.py:47-53), which needs weight gradients inside the backbone — calling it with autograd
enabled raises instead of silently returning a tensor without a graph.  Tasks: the four the reference's
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
        if backbone is None:
            from transformers.models.gpt2.modeling_gpt2 import GPT2Model
            hf = GPT2Model.from_pretrained("gpt2", output_attentions=True, output_hidden_states=True)
            hf.h = hf.h[: self.gpt_layers]
            hf.config.n_layer = len(hf.h)
            object.__setattr__(self, "_hf_model", hf)
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
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise MtsError("GPT4TS: the training path (LayerNorm / wpe gradients inside the GPT-2 blocks) is not "
                           "implemented on the kernel stack; run inference under torch.no_grad()")
        if x.dtype != torch.float32:
            raise MtsError(f"x_enc must be fp32, got {x.dtype}")
        x = x.contiguous()
        B, T, C = x.shape
        assert T == self.seq_len and C == self.enc_in
        if not self.use_cuda_graph:
            return self._forward_impl(x)
        key = (tuple(x.shape), x.device.index, self.training, tuple((p._version, p.data_ptr()) for p in self.parameters()))
        return self._graph.run(key, x, self._forward_impl)

    def _forward_impl(self, x):
        B, T, C = x.shape
        bb = self._backbone
        D, dev = bb.spec.hidden, x.device
        if self.d_model > D or self.d_ff > D or C > D:
            raise MtsError(f"d_model / d_ff / n_features must not exceed the GPT-2 width {D}")
        T2 = T + self.pred_len
        if self.task == "anomaly_detection":
            return self._anomaly_detection(x)
        mean = torch.empty(B, C, device=dev, dtype=torch.float32)
        std = torch.empty(B, C, device=dev, dtype=torch.float32)
        w_conv = self.enc_embedding.value_embedding.tokenConv.weight.detach()
        pe = self.enc_embedding.position_embedding.pe[0, :T]
        X = torch.empty(B * T2, D, device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream().cuda_stream
        if self.task == "forecasting":
            ld_t = ops.ceil8(T)
            enc_t = torch.empty(B, self.d_model, ld_t, device=dev, dtype=torch.bfloat16)
            _lib.call("mts_gpt4ts_embed", x.data_ptr(), w_conv.data_ptr(), pe.data_ptr(), 0, mean.data_ptr(),
                      std.data_ptr(), enc_t.data_ptr(), 0, B, T, C, self.d_model, D, ld_t, 0, 1e-5, stream)
            # rows of X = wpe[t] (+ 0): HF's GPT2Model adds the position table to inputs_embeds (modeling_gpt2.py:584-585)
            ops.prompt_gather(None, bb.embed, bb.wpe, X, rep=1, Lp=0, L=T2, B=B)
            # Linear along time (models/gpt4ts.py:137): X[b, t', :d_model] += W_pre[t', :] . enc[b, :, :] + b_pre[t']
            w_pre = self._bf16_weight("pre", self.predict_linear_pre.weight)          # [T2, ceil8(T)]
            ops.gemm(w_pre, enc_t, X, m=T2, n=self.d_model, k=T, batch=B, lda=w_pre.shape[1], a_bs=0, ldb=ld_t,
                     b_bs=self.d_model * ld_t, ldd=D, d_bs=T2 * D, bias=self.predict_linear_pre.bias.detach(),
                     bias_axis=BIAS_M, epilogue=EPI_RESID_ADD)
        else:
            _lib.call("mts_gpt4ts_embed", x.data_ptr(), w_conv.data_ptr(), pe.data_ptr(), bb.wpe.data_ptr(),
                      mean.data_ptr(), std.data_ptr(), 0, X.data_ptr(), B, T, C, self.d_model, D, 0, 1, 1e-5, stream)
        out = self._blocks_and_out_layer(X, B, T2)
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

    def _blocks_and_out_layer(self, X, B, T2):
        """GPT-2 blocks + ln_f, then out_layer on the first d_ff features (models/gpt4ts.py:140-143): fp32 [B, T2, n_out]."""
        bb = self._backbone
        D = bb.spec.hidden
        hid, _ = bb.forward(X, B, T2)                                                  # bf16 [B*T2, D], ln_f applied
        n_out = self.out_layer.weight.shape[0]
        w_out = self._bf16_weight("out", self.out_layer.weight)                        # [n_out, ceil8(d_ff)]
        out = torch.empty(B * T2, n_out, device=X.device, dtype=torch.float32)
        ops.gemm(hid, w_out, out, m=B * T2, n=n_out, k=self.d_ff, lda=D, ldb=w_out.shape[1],
                 bias=self.out_layer.bias.detach(), bias_axis=BIAS_N)
        return out.view(B, T2, n_out)

    def _anomaly_detection(self, x):
        """models/gpt4ts.py:151-177.  The statistics are taken over "segments" of seg_num = 1 time step (:155-159):
        mean = x, the centred series is identically zero, stdev = sqrt(1e-5).  The GPT-2 therefore sees the same
        input for every sample — zeros plus its position table — and the prediction is dec * sqrt(1e-5) + x
        (:172-175).  That one sequence goes through the blocks once; the de-normalisation broadcasts it."""
        B, T, C = x.shape
        bb = self._backbone
        X = torch.empty(T, bb.spec.hidden, device=x.device, dtype=torch.float32)
        ops.prompt_gather(None, bb.embed, bb.wpe, X, rep=1, Lp=0, L=T, B=1)            # 0 + wpe[t]
        dec = self._blocks_and_out_layer(X, 1, T)                                      # [1, T, C]
        out = dec.expand(B, T, C).contiguous()
        std = torch.full((1, B * T * C), float(torch.tensor(1e-5, dtype=torch.float32).sqrt()), device=x.device)
        ops.revin_denorm(out.view(1, 1, -1), x.reshape(1, -1), std)                    # out * sqrt(1e-5) + x
        return out
