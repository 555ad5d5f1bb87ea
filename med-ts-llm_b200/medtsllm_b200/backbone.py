"""Frozen LLM backbones (Llama / GPT-2 decoder stacks) laid out for the sm_100a kernels.

Replaces the HuggingFace backbone call `self.llm(inputs_embeds=...)` (models/medtsllm.py:346-351):
the arithmetic of HF:models/llama/modeling_llama.py:303-332,375-425 and
HF:models/gpt2/modeling_gpt2.py:262-309,522-636 is re-implemented on libmtsb200 kernels.

HBM layout (per layer, all K-major so every GEMM is the one NT tcgen05 kernel):
  Llama: wqkv [3D,D] bf16 (q|k|v rows), wo [D,D], wgu packed [2*Ipad,D] (128 gate rows then the
         matching 128 up rows, for the fused SwiGLU epilogue), wdown [D,Ipad]; RMSNorm weights fp32.
  GPT-2: Conv1D weights are [in,out]; stored transposed: wqkv [3D,D], wo [D,D], wfc [4D,D],
         wproj [D,4D]; biases and LayerNorm parameters fp32; wpe fp32.
  For the training path each frozen weight additionally keeps its transpose (dgrad is then the same
  NT kernel): ~2x weight memory (26 GB for Llama-2-7B), sized for the 180 GB of a B200.
The residual stream is fp32 [B', L, D] (as in the reference under bf16 autocast, where the residual
adds promote to fp32); GEMM operands are bf16, accumulation fp32 in TMEM.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops
from ._lib import BIAS_N, BIAS_NONE, EPI_GELU_NEW, EPI_RESID_ADD, EPI_ROPE_QK, EPI_SWIGLU, MtsError


@dataclass
class BackboneSpec:
    kind: str            # "llama" | "gpt2"
    hidden: int
    layers: int
    heads: int
    inter: int
    vocab: int
    eps: float
    rope_theta: float = 10000.0
    max_pos: int = 4096
    kv_heads: int = 0    # grouped-query checkpoints: K / V heads of the checkpoint (0 = heads).  The kernels see `heads` K / V
                         # heads: from_hf repeats each K / V head's projection rows heads / kv_heads times, which is exactly
                         # HuggingFace's repeat_kv (HF:models/llama/modeling_llama.py:186-196) folded into the weights

    @property
    def head_dim(self):
        return self.hidden // self.heads


def spec_from_hf_config(cfg) -> BackboneSpec:
    mt = cfg.model_type
    if mt == "llama":
        kv = getattr(cfg, "num_key_value_heads", None) or cfg.num_attention_heads
        if cfg.num_attention_heads % kv or getattr(cfg, "head_dim", None) not in (None, cfg.hidden_size // cfg.num_attention_heads):
            raise MtsError("unsupported attention geometry (K / V heads must divide the heads, head_dim = hidden / heads)")
        rp = getattr(cfg, "rope_parameters", None) or {}
        theta = rp.get("rope_theta", getattr(cfg, "rope_theta", 10000.0))
        if rp.get("rope_type", "default") != "default":
            raise MtsError(f"rope_type {rp.get('rope_type')} is not supported")
        return BackboneSpec("llama", cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                            cfg.intermediate_size, cfg.vocab_size, cfg.rms_norm_eps, theta,
                            cfg.max_position_embeddings, kv_heads=0 if kv == cfg.num_attention_heads else kv)
    if mt == "gpt2":
        inner = cfg.n_inner if cfg.n_inner is not None else 4 * cfg.n_embd
        if cfg.activation_function != "gelu_new":
            raise MtsError(f"GPT-2 activation {cfg.activation_function} is not supported")
        return BackboneSpec("gpt2", cfg.n_embd, cfg.n_layer, cfg.n_head, inner, cfg.vocab_size,
                            cfg.layer_norm_epsilon, max_pos=cfg.n_positions)
    raise MtsError(f"backbone model_type {mt!r} is not supported (llama, gpt2)")


def _rope_tables(L, hd, theta, device):
    # HF:models/llama/modeling_llama.py:107-136 (default rope), fp32
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    f = torch.arange(L, dtype=torch.float32)[:, None] * inv[None, :]
    return f.cos().contiguous().to(device), f.sin().contiguous().to(device)


class KernelBackbone:
    """Device-resident frozen decoder stack.  Not an nn.Module on purpose: `.to(dtype)` from the
    Trainer (tasks/base.py:41) must not re-cast the bf16 kernel weights."""

    def __init__(self, spec: BackboneSpec, device, precision: str = "bf16"):
        """`precision`: which operand copies of the frozen weights to keep — "bf16" (default path: bf16 operands, fp32
        accumulation), "tf32" (evaluation parity mode, the reference's own evaluation regime: fp32 weights rounded to TF32,
        tcgen05 kind::tf32; the bf16 copies are kept too, training stays on them) or "fp32" (same, plus the low
        TF32 pieces for the 3xTF32 fp32-grade contraction)."""
        if precision not in ("bf16", "tf32", "fp32"):
            raise MtsError(f"precision {precision!r}: expected 'bf16', 'tf32' or 'fp32'")
        self.precision = precision
        self.spec = spec
        self.device = torch.device(device)
        self.layers: list[dict] = []
        self.final_norm_w = None
        self.final_norm_b = None
        self.embed = None       # fp32 [V, D] input embedding table (prompt gather)
        self.embed_t = None     # bf16 [D, V] (K-major operand of the mapping GEMM)
        self.wpe = None         # GPT-2 only, fp32 [max_pos, D]
        self._rope = None
        self.cache_gen = 0      # bumped when a lazily built device table (RoPE) is replaced: captured graphs point into it
        self.i_pad = ((spec.inter + 127) // 128) * 128 if spec.kind == "llama" else spec.inter

    # ---------------------------------------------------------------------------------- building
    def _bf16(self, w: torch.Tensor) -> torch.Tensor:
        """fp32 (any device) -> bf16 on the target device, cast by our kernel."""
        return ops.cast_bf16(w.detach().to(self.device, torch.float32))

    def _t_bf16(self, w: torch.Tensor) -> torch.Tensor:
        return ops.transpose_to_bf16(w.detach().to(self.device, torch.float32))

    def _f32(self, w):
        return w.detach().to(self.device, torch.float32).contiguous()

    # ---- evaluation parity modes: fp32 operands of the kind::tf32 GEMM
    def operand(self, x: torch.Tensor, inplace: bool = False):
        """fp32 tensor -> GEMM operand pair (hi, lo): "tf32" = (nearest TF32, None); "fp32" = the 3xTF32 split."""
        if self.precision == "fp32":
            return ops.split_tf32(x)
        return (ops.round_tf32(x, out=x if inplace else None), None)

    def _w32(self, w: torch.Tensor):
        """Frozen fp32 weight [n, k] (any device) -> operand pair on the target device."""
        t = w.detach().to(self.device, torch.float32).contiguous()
        if t.data_ptr() == w.data_ptr():
            t = t.clone()                  # never round the caller's tensor in place
        return self.operand(t, inplace=True)

    @property
    def precise(self) -> bool:
        return self.precision != "bf16"

    def gemm32(self, a_op, w_op, d, **kw):
        """mts_gemm on fp32 operand pairs (kind::tf32; 3xTF32 when both pairs carry low pieces)."""
        if a_op[1] is None or w_op[1] is None:      # "tf32" on a stack built for "fp32": the hi pieces ARE the TF32 operands
            return ops.gemm(a_op[0], w_op[0], d, **kw)
        return ops.gemm(a_op[0], w_op[0], d, a_lo=a_op[1], b_lo=w_op[1], **kw)

    def set_precision(self, precision: str):
        """Switch the evaluation regime of an already built stack: "bf16" always works; "tf32" needs fp32 operands
        (built with "tf32" or "fp32": the hi piece of the 3xTF32 split is the nearest-TF32 weight); "fp32" needs the split."""
        have = getattr(self, "_built_precision", self.precision)
        ok = {"bf16": ("bf16",), "tf32": ("bf16", "tf32"), "fp32": ("bf16", "tf32", "fp32")}[have]
        if precision not in ok:
            raise MtsError(f"this backbone was built for {have!r}: cannot switch to {precision!r}")
        self._built_precision = have
        self.precision = precision
        self.cache_gen += 1

    def embed_t_f32(self):
        """fp32 [D, ceil4(V)] operand pair of the mapping GEMM (built on first use: 0.5 GB for Llama-2-7B)."""
        cache = self.__dict__.setdefault("_embed_t_f32", {})
        if self.precision not in cache:
            V, D = self.embed.shape
            t = torch.zeros(D, (V + 3) // 4 * 4, device=self.device, dtype=torch.float32)
            t[:, :V] = self.embed.t()
            cache[self.precision] = self.operand(t, inplace=True)
        return cache[self.precision]

    def _finish_embeddings(self, emb: torch.Tensor):
        self.embed = self._f32(emb)
        V, D = self.embed.shape
        # [D, ceil8(V)] zero padded: GPT-2's V = 50257 is odd and TMA rows must be 16-byte aligned
        self.embed_t = ops.transpose_strided(self.embed, rows=V, cols=D)

    def _add_llama_layer(self, wq, wk, wv, wo, wg, wu, wd, ln1, ln2):
        s = self.spec
        lay = {}
        lay["wqkv"] = self._bf16(torch.cat([wq, wk, wv], dim=0))
        lay["wo"] = self._bf16(wo)
        lay["wgu"] = ops.pack_gate_up(self._bf16(wg), self._bf16(wu))
        wd_b = self._bf16(wd)                                    # [D, I]
        if self.i_pad != s.inter:
            pad = torch.zeros(s.hidden, self.i_pad, device=self.device, dtype=torch.bfloat16)
            pad[:, : s.inter] = wd_b
            wd_b = pad
        lay["wdown"] = wd_b
        lay["ln1"] = self._f32(ln1)
        lay["ln2"] = self._f32(ln2)
        if self.precise:
            D, I, Ip = s.hidden, s.inter, self.i_pad
            lay["wqkv_f32"] = self._w32(torch.cat([wq, wk, wv], dim=0))
            lay["wo_f32"] = self._w32(wo)
            gu = torch.zeros(2, Ip, D, device=self.device, dtype=torch.float32)
            gu[0, :I], gu[1, :I] = wg.to(self.device, torch.float32), wu.to(self.device, torch.float32)
            # same packing as mts_pack_gate_up: blocks of 128 gate rows followed by the matching 128 up rows
            lay["wgu_f32"] = self._w32(gu.view(2, Ip // 128, 128, D).permute(1, 0, 2, 3).reshape(2 * Ip, D))
            wd32 = torch.zeros(D, Ip, device=self.device, dtype=torch.float32)
            wd32[:, :I] = wd.to(self.device, torch.float32)
            lay["wdown_f32"] = self._w32(wd32)
        self.layers.append(lay)

    def _add_gpt2_layer(self, c_attn_w, c_attn_b, c_proj_w, c_proj_b, fc_w, fc_b, proj_w, proj_b,
                        ln1w, ln1b, ln2w, ln2b):
        lay = {}
        lay["wqkv"] = self._t_bf16(c_attn_w)       # Conv1D [D,3D] -> [3D,D]
        lay["bqkv"] = self._f32(c_attn_b)
        lay["wo"] = self._t_bf16(c_proj_w)
        lay["bo"] = self._f32(c_proj_b)
        lay["wfc"] = self._t_bf16(fc_w)            # [D,4D] -> [4D,D]
        lay["bfc"] = self._f32(fc_b)
        lay["wproj"] = self._t_bf16(proj_w)        # [4D,D] -> [D,4D]
        lay["bproj"] = self._f32(proj_b)
        lay["ln1"], lay["ln1b"] = self._f32(ln1w), self._f32(ln1b)
        lay["ln2"], lay["ln2b"] = self._f32(ln2w), self._f32(ln2b)
        if self.precise:           # Conv1D weights are [in, out]: K-major operands are their transposes
            for name, w in (("wqkv", c_attn_w), ("wo", c_proj_w), ("wfc", fc_w), ("wproj", proj_w)):
                lay[name + "_f32"] = self._w32(w.to(self.device, torch.float32).t())
        self.layers.append(lay)

    @classmethod
    def from_hf(cls, hf_model, device, precision: str = "bf16") -> "KernelBackbone":
        """Converts a HuggingFace LlamaModel / GPT2Model (as loaded by the reference's setup_llm,
        models/medtsllm.py:175-185) layer by layer."""
        spec = spec_from_hf_config(hf_model.config)
        self = cls(spec, device, precision=precision)
        sd = hf_model.state_dict()
        if spec.kind == "llama":
            hd, rep = spec.head_dim, (spec.heads // spec.kv_heads if spec.kv_heads else 1)

            def kv_rows(w):       # [kv_heads*hd, D] -> [heads*hd, D]: query head h reads K / V head h // rep
                return w if rep == 1 else w.view(-1, hd, w.shape[1]).repeat_interleave(rep, dim=0).reshape(-1, w.shape[1])
            for i in range(spec.layers):
                p = f"layers.{i}."
                self._add_llama_layer(
                    sd[p + "self_attn.q_proj.weight"], kv_rows(sd[p + "self_attn.k_proj.weight"]),
                    kv_rows(sd[p + "self_attn.v_proj.weight"]), sd[p + "self_attn.o_proj.weight"],
                    sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"],
                    sd[p + "mlp.down_proj.weight"], sd[p + "input_layernorm.weight"],
                    sd[p + "post_attention_layernorm.weight"])
            self.final_norm_w = self._f32(sd["norm.weight"])
            self._finish_embeddings(sd["embed_tokens.weight"])
        else:
            for i in range(spec.layers):
                p = f"h.{i}."
                self._add_gpt2_layer(
                    sd[p + "attn.c_attn.weight"], sd[p + "attn.c_attn.bias"],
                    sd[p + "attn.c_proj.weight"], sd[p + "attn.c_proj.bias"],
                    sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"],
                    sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"],
                    sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
            self.final_norm_w, self.final_norm_b = self._f32(sd["ln_f.weight"]), self._f32(sd["ln_f.bias"])
            self.wpe = self._f32(sd["wpe.weight"])
            self._finish_embeddings(sd["wte.weight"])
        return self

    @classmethod
    def random_init(cls, spec: BackboneSpec, device, seed=0, std=0.02, precision="bf16"):
        """Seeded random-init stack generated directly on the device (no checkpoints exist offline;
        HF `initializer_range` = 0.02, norms = 1, biases = 0).  `precision`: "bf16" | "tf32" | "fp32" (see __init__)."""
        self = cls(spec, device, precision=precision)
        g = torch.Generator(device=self.device).manual_seed(seed)
        D, I = spec.hidden, spec.inter

        def rnd(*shape):
            return torch.randn(*shape, device=self.device, dtype=torch.float32, generator=g) * std

        ones, zeros = (lambda n: torch.ones(n, device=self.device)), (lambda n: torch.zeros(n, device=self.device))
        for _ in range(spec.layers):
            if spec.kind == "llama":
                self._add_llama_layer(rnd(D, D), rnd(D, D), rnd(D, D), rnd(D, D), rnd(I, D), rnd(I, D),
                                      rnd(D, I), ones(D), ones(D))
            else:
                self._add_gpt2_layer(rnd(D, 3 * D), zeros(3 * D), rnd(D, D), zeros(D), rnd(D, I), zeros(I),
                                     rnd(I, D), zeros(D), ones(D), zeros(D), ones(D), zeros(D))
        self.final_norm_w = ones(D)
        if spec.kind == "gpt2":
            self.final_norm_b = zeros(D)
            self.wpe = rnd(spec.max_pos, D)
        self._finish_embeddings(rnd(spec.vocab, D))
        return self

    # ---------------------------------------------------------------------------------- forward
    def rope(self, L):
        if self.spec.kind != "llama":
            return None
        if self._rope is None or self._rope[0].shape[0] < L:
            self._rope = _rope_tables(max(L, 512), self.spec.head_dim, self.spec.rope_theta, self.device)
            self.cache_gen += 1
        return self._rope

    def weight_bytes(self) -> int:
        """Bytes of the bf16 operand copies + norm / bias vectors one forward streams (default path)."""
        n = 0
        for lay in self.layers:
            n += sum(t.numel() * t.element_size() for k, t in lay.items() if isinstance(t, torch.Tensor) and not k.endswith("_t"))
        return n

    def ensure_transposed(self):
        """Training needs every frozen weight transposed too (dgrad = the same NT GEMM).  Built once,
        lazily, by our transpose kernel; doubles the weight footprint (see module docstring)."""
        if self.layers and "wqkv_t" in self.layers[0]:
            return
        for lay in self.layers:
            lay["wqkv_t"] = ops.transpose_to_bf16(lay["wqkv"])       # [D, 3D]
            lay["wo_t"] = ops.transpose_to_bf16(lay["wo"])           # [D, D]
            if self.spec.kind == "llama":
                lay["wgu_t"] = ops.transpose_to_bf16(lay["wgu"])     # [D, 2*Ipad] (packed column order)
                lay["wdown_t"] = ops.transpose_to_bf16(lay["wdown"])  # [Ipad, D]
            else:
                lay["wfc_t"] = ops.transpose_to_bf16(lay["wfc"])     # [D, I]
                lay["wproj_t"] = ops.transpose_to_bf16(lay["wproj"])  # [I, D]

    def embed_bf16(self):
        """bf16 [V, D] copy of the embedding table (B operand of dW_map = dSource E^T)."""
        if getattr(self, "_embed_bf16", None) is None:
            self._embed_bf16 = ops.cast_bf16(self.embed)
        return self._embed_bf16

    def forward(self, x: torch.Tensor, Bp: int, L: int, stash: list | None = None, lora=None, Lc: int = 0,
                dropout: dict | None = None):
        """x: fp32 residual stream [Bp*L, D].  Inference (stash None): updated IN PLACE layer by layer.
        Training (stash = list): every residual write goes to a fresh buffer (the GEMM epilogue reads
        C = previous stream, writes D = new one) and the per-layer tensors backward() needs are
        appended to `stash`.  Returns (final-norm output bf16 [Bp*L, D], final residual stream fp32).

        Lc > 0: shared-prefix row layout (include/mts_b200.h, mts_attn_causal_shared) — the Lc leading prompt
        positions every sample shares are carried once: x is [Lc + Bp*(L-Lc), D], rows [0, Lc) the prefix, then
        L-Lc own rows per sample.  All row-wise kernels and GEMMs are layout-agnostic; RoPE positions and
        attention are told about it.

        `dropout` = {"embd": p, "attn": p, "resid": p, "seeds": [1 + 3*layers ints]}: the frozen backbone's OWN dropouts,
        which stay live in the reference's train mode because `model.train()` also flips the HF module
        (tasks/forecasting.py:18): GPT-2 embd_pdrop on inputs_embeds + wpe (HF:models/gpt2/modeling_gpt2.py:612),
        attn_pdrop on the attention probabilities (:67-68), resid_pdrop on both projected branches (:233, :243); Llama
        attention_dropout (HF:models/llama/modeling_llama.py:217).  Counter-based masks (mts_dropout) re-created by
        backward() from the same seeds.  Plain row layout only: the masks differ per sample, prompt rows included."""
        s = self.spec
        D, H, hd = s.hidden, s.heads, s.head_dim
        Ls = L - Lc
        M = Lc + Bp * Ls
        if x.shape != (M, D) or x.dtype != torch.float32 or not x.is_contiguous():
            raise MtsError("backbone.forward expects a contiguous fp32 [Bp*L, D] residual stream")
        if s.kind == "gpt2" and L > s.max_pos:
            raise IndexError(f"sequence length {L} exceeds GPT-2 position table {s.max_pos}")
        rope = self.rope(L)
        dev = x.device
        train = stash is not None
        bf = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.bfloat16)  # noqa: E731
        h, qkv, att = bf(M, D), bf(M, 3 * D), bf(M, D)
        llama = s.kind == "llama"
        p_embd = p_attn = p_resid = 0.0
        if dropout is not None:
            if Lc:
                raise MtsError("backbone dropout needs the plain row layout (masks differ per sample)")
            p_embd, p_attn, p_resid = dropout["embd"], dropout["attn"], dropout["resid"]
            seeds = dropout["seeds"]
            if p_embd > 0:
                ops.dropout(x, p_embd, seeds[0], out=x)
        for li, lay in enumerate(self.layers):
            s_attn, s_r1, s_r2 = seeds[1 + 3 * li: 4 + 3 * li] if dropout is not None else (0, 0, 0)
            if train:
                qkv, att = bf(M, 3 * D), bf(M, D)
                if lora is not None:
                    h = bf(M, D)          # the LoRA backward needs this layer's normed input
            x_in = x
            # --- attention half
            fused_rope = llama and lora is None      # RoPE must follow the LoRA update of q: keep it separate then
            if llama:
                ops.rmsnorm(x_in, lay["ln1"], s.eps, out=h)
                if fused_rope:   # q | k rotated in fp32 straight from the accumulators, in the GEMM epilogue
                    ops.gemm(h, lay["wqkv"], qkv, m=M, n=3 * D, k=D, epilogue=EPI_ROPE_QK, rope=rope, rope_L=Ls,
                             rope_hd=hd, rope_cols=2 * D, rope_prefix=Lc)
                else:
                    ops.gemm(h, lay["wqkv"], qkv, m=M, n=3 * D, k=D)
            else:
                ops.layernorm(x_in, lay["ln1"], lay["ln1b"], s.eps, out=h)
                ops.gemm(h, lay["wqkv"], qkv, m=M, n=3 * D, k=D, bias=lay["bqkv"], bias_axis=BIAS_N)
            lora_t = lora.forward_layer(li, h, qkv, M, D) if lora is not None else None
            h_attn = h
            lse = None
            if rope is not None and not fused_rope:    # q, k rotated in place; attention stages them as is
                if Lc:
                    ops.rope_qk_shared_(qkv, Bp, Lc, Ls, H, hd, rope)
                else:
                    ops.rope_qk_(qkv, Bp, L, H, hd, rope)
            if p_attn > 0:
                _, lse = ops.attn_causal_dropout(qkv, Bp, L, H, hd, p_attn, s_attn, out=att)
            elif Lc:
                if train:
                    _, lse = ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, out=att, want_lse=True)
                else:
                    ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, out=att)
            elif train:
                _, lse = ops.attn_causal(qkv, Bp, L, H, hd, rope=None, out=att, want_lse=True)
            else:
                ops.attn_causal(qkv, Bp, L, H, hd, rope=None, out=att)
            x_mid = torch.empty_like(x_in) if train else x_in
            ops.gemm(att, lay["wo"], x_mid, m=M, n=D, k=D, epilogue=EPI_RESID_ADD, c=x_in if train else None,
                     bias=None if llama else lay["bo"], bias_axis=BIAS_NONE if llama else BIAS_N,
                     drop_p=p_resid, drop_seed=s_r1)
            # --- MLP half
            x_out = torch.empty_like(x_in) if train else x_in
            if train and lora is not None:
                h = bf(M, D)
            if llama:
                ops.rmsnorm(x_mid, lay["ln2"], s.eps, out=h)
                if train:   # the SwiGLU epilogue also keeps the gate/up pre-activations (packed column order)
                    pre = bf(M, 2 * self.i_pad)
                    act = bf(M, self.i_pad)
                    ops.gemm(h, lay["wgu"], act, m=M, n=2 * self.i_pad, k=D, epilogue=EPI_SWIGLU, block_n=256, aux=pre)
                else:
                    pre = None
                    act = bf(M, self.i_pad)
                    ops.gemm(h, lay["wgu"], act, m=M, n=2 * self.i_pad, k=D, epilogue=EPI_SWIGLU, block_n=256)
                ops.gemm(act, lay["wdown"], x_out, m=M, n=D, k=self.i_pad, epilogue=EPI_RESID_ADD,
                         c=x_mid if train else None)
            else:
                ops.layernorm(x_mid, lay["ln2"], lay["ln2b"], s.eps, out=h)
                if train:
                    pre = bf(M, s.inter)
                    ops.gemm(h, lay["wfc"], pre, m=M, n=s.inter, k=D, bias=lay["bfc"], bias_axis=BIAS_N)
                    act = ops.gelu_new(pre)
                else:
                    pre = None
                    act = bf(M, s.inter)
                    ops.gemm(h, lay["wfc"], act, m=M, n=s.inter, k=D, bias=lay["bfc"], bias_axis=BIAS_N,
                             epilogue=EPI_GELU_NEW)
                ops.gemm(act, lay["wproj"], x_out, m=M, n=D, k=s.inter, bias=lay["bproj"], bias_axis=BIAS_N,
                         epilogue=EPI_RESID_ADD, c=x_mid if train else None, drop_p=p_resid, drop_seed=s_r2)
            if train:
                stash.append(dict(x_in=x_in, x_mid=x_mid, qkv=qkv, att=att, lse=lse, pre=pre,
                                  h=h_attn if lora is not None else None, lora_t=lora_t))
            x = x_out
        out = bf(M, D)
        if llama:
            ops.rmsnorm(x, self.final_norm_w, s.eps, out=out)
        else:
            ops.layernorm(x, self.final_norm_w, self.final_norm_b, s.eps, out=out)
        return out, x

    def forward_f32(self, x: torch.Tensor, Bp: int, L: int, Lc: int = 0, hidden: list | None = None, lora=None):
        """Evaluation parity modes ("tf32" / "fp32", see __init__): the same blocks with fp32 activations end to end,
        every contraction on tcgen05 kind::tf32 (3xTF32 in "fp32" mode), fp32 attention.  x: fp32 residual stream
        [Lc + Bp*(L-Lc), D], updated in place; returns the final-norm output fp32 (same rows).  `hidden` (list): receives
        a copy of the residual stream entering every layer plus the final one (tests).  Inference only."""
        if not self.precise:
            raise MtsError("this backbone was built without fp32 operands (precision='bf16')")
        s = self.spec
        D, H, hd = s.hidden, s.heads, s.head_dim
        Ls = L - Lc
        M = Lc + Bp * Ls
        if x.shape != (M, D) or x.dtype != torch.float32 or not x.is_contiguous():
            raise MtsError("backbone.forward_f32 expects a contiguous fp32 [rows, D] residual stream")
        if s.kind == "gpt2" and L > s.max_pos:
            raise IndexError(f"sequence length {L} exceeds GPT-2 position table {s.max_pos}")
        rope = self.rope(L)
        dev = x.device
        f32 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)  # noqa: E731
        h, qkv, att = f32(M, D), f32(M, 3 * D), f32(M, D)
        llama = s.kind == "llama"
        tf32 = self.precision == "tf32"
        for li, lay in enumerate(self.layers):
            if hidden is not None:
                hidden.append(x.clone())
            wqkv = lay["wqkv_f32"] if lora is None else lora.folded_qkv(li, lay["wqkv_f32"], self)
            if llama:
                ops.rmsnorm(x, lay["ln1"], s.eps, out=h)
                self.gemm32(self.operand(h, inplace=True), wqkv, qkv, m=M, n=3 * D, k=D, epilogue=EPI_ROPE_QK,
                            rope=rope, rope_L=Ls, rope_hd=hd, rope_cols=2 * D, rope_prefix=Lc)
            else:
                ops.layernorm(x, lay["ln1"], lay["ln1b"], s.eps, out=h)
                self.gemm32(self.operand(h, inplace=True), wqkv, qkv, m=M, n=3 * D, k=D, bias=lay["bqkv"],
                            bias_axis=BIAS_N)
            # "tf32": Q K^T and P V on the tensor cores in TF32 like the reference's matmuls, output already rounded for the
            # out-projection; "fp32": plain fp32 arithmetic
            ops.attn_causal_f32(qkv, Bp, Lc, Ls, H, hd, out=att, tf32=tf32, round_out=tf32)
            self.gemm32((att, None) if tf32 else self.operand(att), lay["wo_f32"], x, m=M, n=D, k=D, epilogue=EPI_RESID_ADD,
                        bias=None if llama else lay["bo"], bias_axis=BIAS_NONE if llama else BIAS_N)
            if llama:
                ops.rmsnorm(x, lay["ln2"], s.eps, out=h)
                act = f32(M, self.i_pad)
                self.gemm32(self.operand(h, inplace=True), lay["wgu_f32"], act, m=M, n=2 * self.i_pad, k=D,
                            epilogue=EPI_SWIGLU, block_n=256, round_tf32=tf32)
                a_op = (act, None) if tf32 else self.operand(act)
                self.gemm32(a_op, lay["wdown_f32"], x, m=M, n=D, k=self.i_pad, epilogue=EPI_RESID_ADD)
            else:
                ops.layernorm(x, lay["ln2"], lay["ln2b"], s.eps, out=h)
                act = f32(M, s.inter)
                self.gemm32(self.operand(h, inplace=True), lay["wfc_f32"], act, m=M, n=s.inter, k=D, bias=lay["bfc"],
                            bias_axis=BIAS_N, epilogue=EPI_GELU_NEW, round_tf32=tf32)
                a_op = (act, None) if tf32 else self.operand(act)
                self.gemm32(a_op, lay["wproj_f32"], x, m=M, n=D, k=s.inter, bias=lay["bproj"], bias_axis=BIAS_N,
                            epilogue=EPI_RESID_ADD)
            del act, a_op
        if hidden is not None:
            hidden.append(x.clone())
        out = f32(M, D)
        if llama:
            ops.rmsnorm(x, self.final_norm_w, s.eps, out=out)
        else:
            ops.layernorm(x, self.final_norm_w, self.final_norm_b, s.eps, out=out)
        return out

    def backward(self, dhid: torch.Tensor, x_final: torch.Tensor, stash: list, Bp: int, L: int, lora=None,
                 Lc: int = 0, norm_grads: dict | None = None, dropout: dict | None = None):
        """dgrad through the frozen stack: dhid = dL/d(final-norm output) bf16 [Bp*L, D] -> returns
        (dL/d(input residual stream) fp32 [Bp*L, D], LoRA gradients aligned with lora.params() or None).
        No gradients for the frozen weights.

        Lc > 0 (shared-prefix layout): with a frozen backbone the prefix rows have no trainable ancestor,
        so the whole chain runs on the samples' own rows only — dhid and the result are [Bp*(L-Lc), D] and
        every stashed tensor is read from row Lc on.  With LoRA the prefix rows matter (they reach the A/B
        pairs): dhid and the result then hold all Lc + Bp*(L-Lc) rows.

        `norm_grads` (dict): also collect the (weight, bias) gradients of every norm — keys ("ln_f",), (layer, "ln1"),
        (layer, "ln2") — for models that train them (GPT4TS, models/gpt4ts.py:47-53)."""
        s = self.spec
        D, H, hd = s.hidden, s.heads, s.head_dim
        Ls = L - Lc
        # LoRA: the A/B pairs receive gradient through the prefix rows as well -> the chain runs on all rows
        full = Lc > 0 and lora is not None
        M = Lc + Bp * Ls if full else Bp * Ls
        own = slice(0 if full else Lc, None)
        self.ensure_transposed()
        rope = self.rope(L)
        dev = dhid.device
        llama = s.kind == "llama"
        norm_bwd = ops.rmsnorm_bwd if llama else ops.layernorm_bwd
        bf = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.bfloat16)  # noqa: E731
        dR = torch.empty(M, D, device=dev, dtype=torch.float32)
        dRb, dH = bf(M, D), bf(M, D)       # dRb: bf16 copy of dR, refreshed by every norm backward
        if norm_grads is not None:
            norm_grads[("ln_f",)] = ops.norm_wgrad(x_final[own], dhid, s.eps, layernorm=not llama)
        norm_bwd(x_final[own], self.final_norm_w, dhid, dR, s.eps, accumulate=False, dx_bf16=dRb)
        lora_grads = [None] * len(lora.params()) if lora is not None else None
        p_embd = p_attn = p_resid = 0.0
        if dropout is not None:                 # the forward's backbone dropouts (plain layout): same seeds, same masks
            p_embd, p_attn, p_resid = dropout["embd"], dropout["attn"], dropout["resid"]
            seeds = dropout["seeds"]
        for li, lay, st in zip(reversed(range(len(self.layers))), reversed(self.layers), reversed(stash)):
            s_attn, s_r1, s_r2 = seeds[1 + 3 * li: 4 + 3 * li] if dropout is not None else (0, 0, 0)
            # gradient entering a residual branch = dR masked like the branch's output was (resid dropout)
            d_mlp = ops.dropout(dRb, p_resid, s_r2) if p_resid > 0 else dRb
            # --- MLP half: x_out = x_mid + W2 act(W1 norm(x_mid))
            if llama:
                dact = bf(M, self.i_pad)
                ops.gemm(d_mlp, lay["wdown_t"], dact, m=M, n=self.i_pad, k=D)
                dpre = ops.swiglu_bwd(st["pre"][own], dact, self.i_pad, 128)
                ops.gemm(dpre, lay["wgu_t"], dH, m=M, n=D, k=2 * self.i_pad)
            else:
                dact = bf(M, s.inter)
                ops.gemm(d_mlp, lay["wproj_t"], dact, m=M, n=s.inter, k=D)
                dpre = ops.gelu_new(st["pre"][own], dact)
                ops.gemm(dpre, lay["wfc_t"], dH, m=M, n=D, k=s.inter)
            if norm_grads is not None:
                norm_grads[(li, "ln2")] = ops.norm_wgrad(st["x_mid"][own], dH, s.eps, layernorm=not llama)
            norm_bwd(st["x_mid"][own], lay["ln2"], dH, dR, s.eps, accumulate=True, dx_bf16=dRb)
            # --- attention half: x_mid = x_in + Wo attn(Wqkv norm(x_in))
            datt = bf(M, D)
            d_att = ops.dropout(dRb, p_resid, s_r1) if p_resid > 0 else dRb
            ops.gemm(d_att, lay["wo_t"], datt, m=M, n=D, k=D)
            if p_attn > 0:
                dqkv = ops.attn_causal_dropout_bwd(st["qkv"], st["att"], datt, st["lse"], Bp, L, H, hd, p_attn, s_attn,
                                                   rope=rope)
            elif full:
                dqkv = ops.attn_causal_shared_bwd_full(st["qkv"], st["att"], datt, st["lse"], Bp, Lc, Ls, H, hd, rope=rope)
            elif Lc:
                dqkv = ops.attn_causal_shared_bwd(st["qkv"], st["att"][own], datt,
                                                  ops.lse_own_view(st["lse"], Bp, Lc, Ls, H), Bp, Lc, Ls, H, hd, rope=rope)
            else:
                dqkv = ops.attn_causal_bwd(st["qkv"], st["att"], datt, st["lse"], Bp, L, H, hd, rope=rope,
                                           pre_roped=rope is not None)
            ops.gemm(dqkv, lay["wqkv_t"], dH, m=M, n=D, k=3 * D)
            if lora is not None:
                dAs, dBs = lora.backward_layer(li, st["h"], st["lora_t"], dqkv, dH, M, D)
                nA = len(lora.A)
                for t in range(len(lora.targets)):
                    lora_grads[lora.index(li, t)] = dAs[t]
                    lora_grads[nA + lora.index(li, t)] = dBs[t]
            if norm_grads is not None:
                norm_grads[(li, "ln1")] = ops.norm_wgrad(st["x_in"][own], dH, s.eps, layernorm=not llama)
            norm_bwd(st["x_in"][own], lay["ln1"], dH, dR, s.eps, accumulate=True, dx_bf16=dRb)
        if p_embd > 0:
            ops.dropout(dR, p_embd, seeds[0], out=dR)
        return dR, lora_grads

    def flops_per_token_fwd(self, L: int) -> float:
        """Dense algorithmic forward FLOPs per token: 2*W_blk + 4*L*D per layer."""
        s = self.spec
        w_blk = 4 * s.hidden ** 2 + (3 if s.kind == "llama" else 2) * s.hidden * s.inter
        return s.layers * (2.0 * w_blk + 4.0 * L * s.hidden)
