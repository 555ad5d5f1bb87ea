"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch fp32) of the reference's GPT4TS forward
(models/gpt4ts.py), the model of BASELINE.json configs[0] (ETTh1 forecasting on a frozen GPT-2).

NOT part of the product: only tests/ (and bench.py's CPU legs) may import this module, as the checker.

Parity status: PINNED.  tests/test_oracle.py checks this restatement against golden tensors produced by
running the unmodified reference class in the build container (oracle/make_golden_gpt4ts.py ->
tests/golden/gpt4ts_*.pt).  The GPT-2 stack itself is `medtsllm_oracle.gpt2_forward`, pinned against
HuggingFace's GPT2Model in the same test file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .medtsllm_oracle import gpt2_forward


def positional_embedding(T: int, d_model: int) -> torch.Tensor:
    """models/layers/embed.py:8-27 — fixed sinusoid table, rows [0, T)."""
    position = torch.arange(0, T).float().unsqueeze(1)
    div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
    pe = torch.zeros(T, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def instance_norm(x: torch.Tensor, eps: float = 1e-5):
    """models/gpt4ts.py:129-133 ("Normalization from Non-stationary Transformer"): per (sample, channel)
    mean and sqrt(biased var + eps) over time.  x [B, T, C] -> (normalised, mean [B,1,C], stdev [B,1,C])."""
    mean = x.mean(1, keepdim=True)
    xc = x - mean
    stdev = torch.sqrt(torch.var(xc, dim=1, keepdim=True, unbiased=False) + eps)
    return xc / stdev, mean, stdev


def data_embedding(xn: torch.Tensor, w_conv: torch.Tensor) -> torch.Tensor:
    """DataEmbedding with x_mark = None, dropout 0 (models/layers/embed.py:109-131): TokenEmbedding =
    Conv1d(C -> d_model, k=3, circular over TIME, bias=False) (:29-46) + the sinusoid table."""
    prev, nxt = torch.roll(xn, 1, dims=1), torch.roll(xn, -1, dims=1)
    val = prev @ w_conv[:, :, 0].T + xn @ w_conv[:, :, 1].T + nxt @ w_conv[:, :, 2].T
    return val + positional_embedding(xn.shape[1], w_conv.shape[0])[None]


def gpt4ts_forward(x_enc, params, gpt2_sd, spec, *, training: bool = False, return_stages: bool = False,
                   dropout_masks=None):
    """GPT4TS.forward (models/gpt4ts.py:83-104) for the tasks the reference's Trainers run it on.

    params: the module's own tensors by reference name (enc_embedding.value_embedding.tokenConv.weight,
            predict_linear_pre.{weight,bias}, out_layer.{weight,bias}).
    spec:   task, pred_len, d_ff, gpt_layers, n_heads, eps, n_classes, seg_mode.
    dropout_masks (train mode): {"embed": multiplicative mask [B, T, d_model] of DataEmbedding.dropout
            (models/layers/embed.py:131) or None, "backbone": the GPT-2 masks of medtsllm_oracle.gpt2_forward or None}."""
    B, T, C = x_enc.shape
    task = spec["task"]
    stages = {}
    if task not in ("forecasting", "anomaly_detection", "semantic_segmentation", "segmentation"):
        raise ValueError("Task name is not valid")          # models/gpt4ts.py:103-104 (e.g. "reconstruction")
    D = gpt2_sd["wpe.weight"].shape[1]
    if task == "anomaly_detection":
        # :151-164.  The "segments" the statistics are taken over have length seg_num = 1 (:155-159), i.e. every
        # time step is its own segment: mean = x, the centred series is identically ZERO, stdev = sqrt(0 + 1e-5).
        # The GPT-2 therefore sees only its position table and the prediction is dec * sqrt(1e-5) + x (:172-175).
        mean = x_enc
        xc = x_enc - mean
        stdev = torch.sqrt(torch.var(xc.unsqueeze(2), dim=2, unbiased=False) + 1e-5)
        enc = xc / stdev
    else:
        xn, mean, stdev = instance_norm(x_enc)
        enc = data_embedding(xn, params["enc_embedding.value_embedding.tokenConv.weight"])
        if dropout_masks is not None and dropout_masks.get("embed") is not None:
            enc = enc * dropout_masks["embed"]
        if task == "forecasting":                            # :137: Linear along time, T -> T + pred
            enc = F.linear(enc.permute(0, 2, 1), params["predict_linear_pre.weight"],
                           params["predict_linear_pre.bias"]).permute(0, 2, 1)
    stages["embedding"] = enc
    enc = F.pad(enc, (0, D - enc.shape[-1]))                 # :138, :164, :212, :242
    dec = gpt2_forward(enc, gpt2_sd, n_layers=spec["gpt_layers"], n_heads=spec["n_heads"], eps=spec.get("eps", 1e-5),
                       dropout=dropout_masks.get("backbone") if dropout_masks is not None else None)
    stages["gpt2"] = dec
    dec = F.linear(dec[:, :, : spec["d_ff"]], params["out_layer.weight"], params["out_layer.bias"])
    if task in ("forecasting", "anomaly_detection"):         # de-normalisation, :146-147, :172-175
        dec = dec * stdev + mean
        if task == "forecasting":
            dec = dec[:, -spec["pred_len"]:, :]              # :91
    else:
        dec = dec.squeeze(-1)
        if not training:
            if task == "semantic_segmentation":              # :221-226
                dec = F.softmax(dec.reshape(B, T, -1), dim=-1) if spec.get("n_classes", 0) > 2 else torch.sigmoid(dec)
            elif spec.get("seg_mode") == "boundary-prediction":   # :251-252
                dec = torch.sigmoid(dec)
    stages["output"] = dec
    return (dec, stages) if return_stages else dec
