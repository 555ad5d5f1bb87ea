"""LoRA adapters on the frozen backbone (BASELINE config 5; models/medtsllm.py:187-204).

The reference wraps the HF model with `peft.get_peft_model(llm, LoraConfig(task_type=FEATURE_EXTRACTION,
r, lora_alpha, init_lora_weights, lora_dropout, use_rslora))` and `layers == "auto"`, i.e. peft's
default target modules: `q_proj`, `v_proj` for Llama, the fused `c_attn` for GPT-2.  `peft` is not
installed in this image and is in neither requirements file of the reference, so this is a restatement
of its published arithmetic (PARITY UNPINNED against peft itself; pinned against
oracle/medtsllm_oracle.py's plain-PyTorch LoRA and by identity-at-init):

    y = W x + scale * B (A x),   scale = alpha / sqrt(r) (rsLoRA, the reference default) or alpha / r,
    A: [r, in] kaiming-uniform(a = sqrt 5),  B: [out, r] zeros  ->  identity at initialisation.

Kernel mapping: the LoRA path stays SEPARATE from the frozen bf16 weight (folding B·A into a bf16 W would
round small updates away): T = h A^T is one skinny tcgen05 GEMM, and q / v (or the whole qkv for GPT-2)
accumulate scale * T B^T through the GEMM's bf16 read-modify-write epilogue.  Backward: dB = s dY^T T,
dA = s (dY B)^T h, dh += s (dY B) A — all mts_gemm on transposed operands.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from ._lib import EPI_RESID_ADD, MtsError


class LoraAdapters(nn.Module):
    """Trainable A/B pairs, registered under `model.llm` like peft's wrapped model so that
    `MedTsLLM.state_dict()` drops them exactly as the reference drops `llm.*` (models/medtsllm.py:238-241)."""

    def __init__(self, spec, rank: int, alpha: float, rslora: bool = True, init: bool = True):
        super().__init__()
        if rank <= 0:
            raise MtsError("lora.rank must be positive")
        self.kind, self.rank = spec.kind, rank
        self.rp = (rank + 7) // 8 * 8                      # padded rank: TMA rows are 16-byte multiples
        self.scale = alpha / math.sqrt(rank) if rslora else alpha / rank
        self.config = None
        D = spec.hidden
        self.targets = ("q", "v") if spec.kind == "llama" else ("c_attn",)
        out_dim = D if spec.kind == "llama" else 3 * D
        self.A = nn.ParameterList()
        self.B = nn.ParameterList()
        for _ in range(spec.layers):
            for _t in self.targets:
                a = torch.empty(rank, D)
                nn.init.kaiming_uniform_(a, a=math.sqrt(5))
                b = torch.zeros(out_dim, rank)
                if not init:
                    nn.init.normal_(b, std=0.02)
                self.A.append(nn.Parameter(a))
                self.B.append(nn.Parameter(b))
        self._cache = {}
        self._fold_cache = {}           # evaluation parity modes: per-layer operand pair of W_qkv + scale * B A
        self._opt_steps = 0             # bumped by the model's global optimizer-step hook (see model._install_optimizer_hook)
        self._cache_gen = 0             # bumped whenever a cached operand set is replaced (graph-replay key)
        self._force_recast = False      # graph capture of a training step: cast even if the cache would hit

    def params(self):
        return list(self.A) + list(self.B)

    def index(self, layer: int, t: int) -> int:
        return layer * len(self.targets) + t

    def device_tensors(self, layer: int):
        """bf16 operands of one layer (re-cast only when the optimizer changed the masters):
        a_cat [T*rp, D] (A of every target stacked, zero rows up to rp), a_cat_t [D, T*rp],
        b[t] [out, rp], b_t[t] [rp, out]."""
        nt = len(self.targets)
        ps = [self.A[self.index(layer, t)] for t in range(nt)] + [self.B[self.index(layer, t)] for t in range(nt)]
        key = tuple((p._version, p.data_ptr()) for p in ps) + (self._opt_steps,)
        hit = self._cache.get(layer)
        if hit is not None and hit[0] == key and not self._force_recast:
            return hit[1]
        dev = ps[0].device
        D = ps[0].shape[1]
        a_cat = torch.zeros(nt * self.rp, D, device=dev, dtype=torch.bfloat16)
        for t in range(nt):
            ops.cast_rows(ps[t].detach(), rows=self.rank, cols=D, out=a_cat[t * self.rp:], ld_out=D)
        b, b_t = [], []
        for t in range(nt):
            B = ps[nt + t].detach()
            bb = ops.cast_rows(B, rows=B.shape[0], cols=self.rank, ld_out=self.rp)     # [out, rp] zero padded
            b.append(bb)
            b_t.append(ops.transpose_strided(bb, rows=B.shape[0], cols=self.rp))        # [rp, out]
        d = dict(a_cat=a_cat, a_cat_t=ops.transpose_strided(a_cat, rows=nt * self.rp, cols=D), b=b, b_t=b_t)
        self._cache[layer] = (key, d)
        self._cache_gen += 1
        return d

    def folded_qkv(self, layer: int, wqkv_op, bb):
        """Evaluation parity modes (fp32 activations): the operand pair of W_qkv + scale * B A.  In fp32 the low-rank
        update can be folded into the frozen weight without losing it to rounding (unlike the bf16 path, which keeps the
        LoRA branch separate); B A itself is a 3xTF32 mts_gemm accumulated onto a copy of W.  Cached until A / B change."""
        nt = len(self.targets)
        ps = [self.A[self.index(layer, t)] for t in range(nt)] + [self.B[self.index(layer, t)] for t in range(nt)]
        key = tuple((p._version, p.data_ptr()) for p in ps) + (self._opt_steps, bb.precision)
        hit = self._fold_cache.get(layer)
        if hit is not None and hit[0] == key:
            return hit[1]
        W = wqkv_op[0].clone() if wqkv_op[1] is None else wqkv_op[0] + wqkv_op[1]        # fp32 [3D, D]
        D = W.shape[1]
        for t, (off, width) in enumerate(self.out_slices(D)):
            A, B = ps[t].detach(), ps[nt + t].detach()                                     # [r, D], [width, r]
            r4 = (self.rank + 3) // 4 * 4
            Bp = torch.zeros(width, r4, device=W.device, dtype=torch.float32)
            Bp[:, :self.rank] = B
            At = torch.zeros(D, r4, device=W.device, dtype=torch.float32)
            At[:, :self.rank] = A.t()
            b_hi, b_lo = ops.split_tf32(Bp)
            a_hi, a_lo = ops.split_tf32(At)
            ops.gemm(b_hi, a_hi, W, m=width, n=D, k=self.rank, lda=r4, ldb=r4, ldd=D, d_off=off * D, a_lo=b_lo, b_lo=a_lo,
                     alpha=self.scale, epilogue=EPI_RESID_ADD)
        op = bb.operand(W, inplace=True)
        self._fold_cache[layer] = (key, op)
        self._cache_gen += 1
        return op

    # ------------------------------------------------------------------------------------------
    def out_slices(self, D):
        """(column offset in qkv, width) of each target's output."""
        return [(0, D), (2 * D, D)] if self.kind == "llama" else [(0, 3 * D)]

    def forward_layer(self, layer: int, h, qkv, M: int, D: int):
        """qkv[:, slice_t] += scale * (h A_t^T) B_t^T.  Returns T = h A_cat^T (kept for the backward)."""
        w = self.device_tensors(layer)
        nt, rp = len(self.targets), self.rp
        T = torch.empty(M, nt * rp, device=h.device, dtype=torch.bfloat16)
        ops.gemm(h, w["a_cat"], T, m=M, n=nt * rp, k=D)
        for t, (off, width) in enumerate(self.out_slices(D)):
            ops.gemm(T, w["b"][t], qkv, m=M, n=width, k=rp, lda=nt * rp, a_off=t * rp, ldb=rp, ldd=3 * D, d_off=off,
                     alpha=self.scale, epilogue=EPI_RESID_ADD)
        return T

    def backward_layer(self, layer: int, h, T, dqkv, dH, M: int, D: int):
        """Adds the LoRA contribution to dH (bf16 [M, D], in place) and returns ([dA_t], [dB_t]) fp32."""
        w = self.device_tensors(layer)
        nt, rp, r, s = len(self.targets), self.rp, self.rank, self.scale
        dev = h.device
        Mp = ops.ceil8(M)
        U = torch.empty(M, nt * rp, device=dev, dtype=torch.bfloat16)                   # dY_t B_t
        T_t = ops.transpose_strided(T, rows=M, cols=nt * rp)                            # [nt*rp, Mp]
        dBs = []
        for t, (off, width) in enumerate(self.out_slices(D)):
            ops.gemm(dqkv, w["b_t"][t], U, m=M, n=rp, k=width, lda=3 * D, a_off=off, ldb=w["b_t"][t].shape[1],
                     ldd=nt * rp, d_off=t * rp)
            dy_t = ops.transpose_strided(dqkv, rows=M, cols=width, ld_in=3 * D, in_off=off)   # [width, Mp]
            dB = torch.empty(width, rp, device=dev, dtype=torch.float32)
            ops.gemm(dy_t, T_t, dB, m=width, n=rp, k=M, lda=Mp, ldb=Mp, b_off=t * rp * Mp, alpha=s)
            dBs.append(dB[:, :r].contiguous() if rp != r else dB)
        # dh += s * U A_cat
        ops.gemm(U, w["a_cat_t"], dH, m=M, n=D, k=nt * rp, ldb=w["a_cat_t"].shape[1], alpha=s, epilogue=EPI_RESID_ADD)
        # dA_cat = s * U^T h
        dA = torch.empty(nt * rp, D, device=dev, dtype=torch.float32)
        ops.gemm(ops.transpose_strided(U, rows=M, cols=nt * rp), ops.transpose_strided(h, rows=M, cols=D), dA,
                 m=nt * rp, n=D, k=M, lda=Mp, ldb=Mp, alpha=s)
        dAs = [dA[t * rp:t * rp + r] for t in range(nt)]
        return dAs, dBs

    # ------------------------------------------------------------------------------------------
    def save_pretrained(self, path, *args, **kwargs):
        """loggers/base_logger.py:42-43 calls `model.llm.save_pretrained(<name>-lora.safetensors)`."""
        from safetensors.torch import save_file
        tensors = {}
        nt = len(self.targets)
        for i in range(len(self.A) // nt):
            for t, name in enumerate(self.targets):
                mod = {"q": "self_attn.q_proj", "v": "self_attn.v_proj", "c_attn": "attn.c_attn"}[name]
                prefix = ("base_model.model.layers." if self.kind == "llama" else "base_model.model.h.") + f"{i}.{mod}"
                tensors[f"{prefix}.lora_A.weight"] = self.A[self.index(i, t)].detach().cpu().contiguous()
                tensors[f"{prefix}.lora_B.weight"] = self.B[self.index(i, t)].detach().cpu().contiguous()
        import os
        os.makedirs(os.path.dirname(str(path)) or ".", exist_ok=True)
        save_file(tensors, str(path))

    def load_pretrained(self, path):
        """Inverse of save_pretrained: reads the A/B pairs back from `<name>-lora.safetensors` (the reference saves
        them, loggers/base_logger.py:42-43, but `BaseTask.from_run_id` never restores them, tasks/base.py:283-306 —
        a maintainer calls `trainer.model.llm.load_pretrained(path)` after it).  Strict: every pair must be present
        with the right shape."""
        from safetensors.torch import load_file
        tensors = load_file(str(path))
        nt = len(self.targets)
        with torch.no_grad():
            for i in range(len(self.A) // nt):
                for t, name in enumerate(self.targets):
                    mod = {"q": "self_attn.q_proj", "v": "self_attn.v_proj", "c_attn": "attn.c_attn"}[name]
                    prefix = ("base_model.model.layers." if self.kind == "llama" else "base_model.model.h.") + f"{i}.{mod}"
                    for which, plist in (("lora_A", self.A), ("lora_B", self.B)):
                        src = tensors[f"{prefix}.{which}.weight"]
                        dst = plist[self.index(i, t)]
                        if src.shape != dst.shape:
                            raise ValueError(f"{prefix}.{which}: shape {tuple(src.shape)} != {tuple(dst.shape)}")
                        dst.copy_(src.to(dst.device, dst.dtype))      # in place: bumps ._version -> bf16 caches refresh
        return sorted(tensors)

