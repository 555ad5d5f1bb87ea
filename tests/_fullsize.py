"""Helpers for the FULL-ARCHITECTURE parity tests (tests/test_full_size_parity_gpu.py): build the kernel model at a
BASELINE workload's real backbone (Llama-2-7B shape: D=4096, I=11008, 32 layers, head dim 128; GPT-2-medium: 24 layers)
on a reduced batch, expose the device-resident weights to the CPU oracle one layer at a time, and (yardstick) run
HuggingFace's own LlamaModel / GPT2Model on the same weights on the GPU in the reference's two regimes."""
from __future__ import annotations

import dataclasses
import re

import torch


def build(workload: str, cuda, batch: int, *, precision: str = "bf16", layers: int | None = None, seed: int = 0,
          dropout: float = 0.0):
    from medtsllm_b200.backbone import KernelBackbone
    from medtsllm_b200.model import MedTsLLM
    from medtsllm_b200.synthetic import (WORKLOADS, AttrDict, FixedLengthTokenizer, SyntheticDataset,
                                         experiment_config, make_inputs)
    w = dataclasses.replace(WORKLOADS[workload], B=batch)
    spec = w.backbone if layers is None else dataclasses.replace(w.backbone, layers=layers)
    w = dataclasses.replace(w, backbone=spec)
    bb = KernelBackbone.random_init(spec, cuda, seed=seed, precision=precision)
    torch.manual_seed(0)
    cfg = experiment_config(w)
    cfg["training"]["dropout"] = dropout
    if precision != "bf16":
        cfg["setup"]["dtype"] = "float32"
    model = MedTsLLM(AttrDict(cfg), SyntheticDataset(w), backbone=bb,
                     tokenizer=FixedLengthTokenizer(spec.vocab, w.prompt_len)).to(cuda, torch.float32).eval()
    return w, model, make_inputs(w)


class LazyBackboneState:
    """Read-only mapping with HuggingFace parameter names over a device-resident KernelBackbone: a layer's tensors are
    copied to the host (fp32) when first asked for and the previously touched layer is dropped, so the CPU oracle can walk
    a 26 GB stack with < 2 GB resident.  `keep=True` keeps everything (autograd through the oracle holds the weights)."""

    def __init__(self, bb, keep: bool = False, precision: str = "bf16"):
        self.bb, self.keep, self.precision = bb, keep, precision
        self._layer, self._cache = None, {}
        self._global = {}

    def _w(self, lay, name):
        """The weight as the oracle's arithmetic sees it: what the kernels multiply by (bf16-rounded in bf16 mode,
        fp32/tf32-rounded in the parity mode)."""
        if self.precision == "bf16":
            return lay[name].float().cpu()
        hi, lo = lay[name + "_f32"]                     # operand pair of the kind::tf32 GEMM (lo: 3xTF32 split only)
        return (hi if lo is None else hi + lo).cpu()

    def _load_layer(self, i):
        bb, s = self.bb, self.bb.spec
        lay = bb.layers[i]
        D, I = s.hidden, s.inter
        out = {}
        if s.kind == "llama":
            q, k, v = self._w(lay, "wqkv").split(D, 0)
            out["self_attn.q_proj.weight"], out["self_attn.k_proj.weight"], out["self_attn.v_proj.weight"] = q, k, v
            out["self_attn.o_proj.weight"] = self._w(lay, "wo")
            wgu = self._w(lay, "wgu").view(-1, 2, 128, D)
            out["mlp.gate_proj.weight"] = wgu[:, 0].reshape(-1, D)[:I].contiguous()
            out["mlp.up_proj.weight"] = wgu[:, 1].reshape(-1, D)[:I].contiguous()
            out["mlp.down_proj.weight"] = self._w(lay, "wdown")[:, :I].contiguous()
            out["input_layernorm.weight"] = lay["ln1"].cpu()
            out["post_attention_layernorm.weight"] = lay["ln2"].cpu()
        else:
            # HF Conv1D stores [in, out]; the kernel layout is the transpose
            out["attn.c_attn.weight"] = self._w(lay, "wqkv").t().contiguous()
            out["attn.c_attn.bias"] = lay["bqkv"].cpu()
            out["attn.c_proj.weight"] = self._w(lay, "wo").t().contiguous()
            out["attn.c_proj.bias"] = lay["bo"].cpu()
            out["mlp.c_fc.weight"] = self._w(lay, "wfc").t().contiguous()
            out["mlp.c_fc.bias"] = lay["bfc"].cpu()
            out["mlp.c_proj.weight"] = self._w(lay, "wproj").t().contiguous()
            out["mlp.c_proj.bias"] = lay["bproj"].cpu()
            out["ln_1.weight"], out["ln_1.bias"] = lay["ln1"].cpu(), lay["ln1b"].cpu()
            out["ln_2.weight"], out["ln_2.bias"] = lay["ln2"].cpu(), lay["ln2b"].cpu()
        return out

    def __getitem__(self, key):
        m = re.match(r"(?:layers|h)\.(\d+)\.(.+)", key)
        if m:
            i, name = int(m.group(1)), m.group(2)
            if self.keep:
                if i not in self._cache:
                    self._cache[i] = self._load_layer(i)
                return self._cache[i][name]
            if self._layer != i:
                self._layer, self._cache = i, {i: self._load_layer(i)}
            return self._cache[i][name]
        if key not in self._global:
            bb = self.bb
            table = {"norm.weight": bb.final_norm_w, "ln_f.weight": bb.final_norm_w, "ln_f.bias": bb.final_norm_b,
                     "embed_tokens.weight": bb.embed, "wte.weight": bb.embed, "wpe.weight": bb.wpe}
            self._global[key] = table[key].float().cpu()
        return self._global[key]


def oracle_spec(w, model):
    s = w.backbone
    return dict(task=w.task, pred_len=w.pred, patch_len=16, stride=8, d_model=32, d_ff=w.d_ff, n_heads=8,
                covariate_mode=w.covariate_mode, downsample="linear", n_outputs_per_step=model.n_outputs_per_step,
                backbone=s.kind, n_layers=s.layers, llm_heads=s.heads, eps=s.eps, rope_theta=s.rope_theta, pad_id=2,
                seg_mode="boundary-prediction", n_classes=w.n_classes)


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def hf_backbone_on_gpu(bb, cuda, precision: str = "bf16"):
    """HuggingFace LlamaModel / GPT2Model (the reference's third-party backbone, models/medtsllm.py:175-185) holding the
    SAME weights as the kernel backbone, fp32 on the GPU, eager attention, output_hidden_states."""
    import transformers
    s = bb.spec
    if s.kind == "llama":
        cfg = transformers.LlamaConfig(hidden_size=s.hidden, intermediate_size=s.inter, num_hidden_layers=s.layers,
                                       num_attention_heads=s.heads, num_key_value_heads=s.heads, vocab_size=8,
                                       rms_norm_eps=s.eps, max_position_embeddings=s.max_pos)
        cls = transformers.LlamaModel
    else:
        cfg = transformers.GPT2Config(n_embd=s.hidden, n_layer=s.layers, n_head=s.heads, vocab_size=8,
                                      n_positions=s.max_pos, attn_pdrop=0.0, embd_pdrop=0.0, resid_pdrop=0.0)
        cls = transformers.GPT2Model
    cfg.output_hidden_states = True
    cfg._attn_implementation = "eager"
    with torch.device(cuda):
        hf = cls(cfg)
    hf = hf.to(cuda, torch.float32).eval()
    sd = LazyBackboneState(bb, precision=precision)
    own = hf.state_dict()
    with torch.no_grad():
        for k, t in own.items():
            if k.startswith(("embed_tokens", "wte")) or k.endswith((".attn.bias", ".attn.masked_bias", "inv_freq")):
                continue
            t.copy_(sd[k].to(cuda))
    return hf


def hf_regimes(hf, x, kind):
    """last_hidden_state + hidden_states of the HF backbone on inputs_embeds x (fp32 [B, L, D], pre-wpe for GPT-2) in
    three regimes: true fp32, the reference's eval regime (TF32 matmuls, tasks/base.py:19-22) and its training regime
    (bf16 autocast, tasks/forecasting.py:22)."""
    res = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.get_float32_matmul_precision())
    try:
        for name, tf32, autocast in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.set_float32_matmul_precision("medium" if tf32 else "highest")
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                o = hf(inputs_embeds=x)
            res[name] = ([h.float() for h in o.hidden_states], o.last_hidden_state.float())
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old[0], old[1]
        torch.set_float32_matmul_precision(old[2])
    return res
