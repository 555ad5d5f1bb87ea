"""Synthetic workloads of BASELINE.json's configs: shapes, experiment configs,
deterministic tokenizer with a fixed 128-token prompt, seeded windows.  No datasets or checkpoints
exist offline, so benchmarks and smoke tests run on these."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .backbone import BackboneSpec

LLAMA2_7B = BackboneSpec("llama", hidden=4096, layers=32, heads=32, inter=11008, vocab=32000, eps=1e-5,
                         rope_theta=10000.0, max_pos=4096)
GPT2_MEDIUM = BackboneSpec("gpt2", hidden=1024, layers=24, heads=16, inter=4096, vocab=50257, eps=1e-5, max_pos=1024)
GPT2_SMALL = BackboneSpec("gpt2", hidden=768, layers=12, heads=12, inter=3072, vocab=50257, eps=1e-5, max_pos=1024)
LLAMA_MINI = BackboneSpec("llama", hidden=256, layers=2, heads=2, inter=512, vocab=512, eps=1e-5, max_pos=512)


@dataclass
class Workload:
    name: str
    backbone: BackboneSpec
    task: str
    B: int
    T: int
    pred: int
    C: int
    d_ff: int = 64
    n_classes: int = 0
    covariate_mode: str = "concat"
    description: str = ""
    prompt_len: int = 128
    lora_rank: int = 0          # > 0: LoRA fine-tuning of the backbone's q/v (Llama) or c_attn (GPT-2) projections

    @property
    def n_patches(self):
        return (self.T + 8 - 16) // 8 + 1

    @property
    def seq(self):
        return self.prompt_len + self.n_patches


WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at 1 GPU
    "bidmc_llama2_7b": Workload("bidmc_llama2_7b", LLAMA2_7B, "segmentation", B=32, T=512, pred=512, C=3,
                                description="The BIDMC dataset contains PPG, ECG and respiration signals."),
    # configs[2]
    # (BASELINE.json names n_vars = 12; the shipped ludb.toml reads one lead at a time with covariate_mode "univariate" —
    # SURVEY.md section 8 sizes this config as 12 leads concatenated: d_model' = 384, 4 x 1024 outputs per window)
    "ludb_llama2_7b": Workload("ludb_llama2_7b", LLAMA2_7B, "semantic_segmentation", B=16, T=1024, pred=1024, C=12,
                               d_ff=128, n_classes=4, covariate_mode="concat",
                               description="LUDB is an ECG signal database with marked boundaries of waves."),
    # configs[3]
    "psm_gpt2_medium": Workload("psm_gpt2_medium", GPT2_MEDIUM, "anomaly_detection", B=64, T=100, pred=100, C=25,
                                description="PSM is a server machine dataset from eBay."),
    # configs[4]: LoRA fine-tune (rank 8, alpha 16, rsLoRA; the shipped configs carry no LoRA block of their own)
    "ventilator_llama2_7b": Workload("ventilator_llama2_7b", LLAMA2_7B, "forecasting", B=16, T=336, pred=96, C=2,
                                     description="Ventilator pressure and flow waveforms.", lora_rank=8),
    # configs[0]'s shape for medtsllm_b200.GPT4TS (bench.py --workload etth1_gpt4ts): GPT-2-small, 6 blocks
    "etth1_gpt4ts": Workload("etth1_gpt4ts", GPT2_SMALL, "forecasting", B=8, T=96, pred=96, C=7, d_ff=768,
                             description="ETTh1", prompt_len=0),
    # tiny shape for smoke tests / CPU-side checks
    "mini_llama": Workload("mini_llama", LLAMA_MINI, "forecasting", B=4, T=96, pred=24, C=3, prompt_len=32,
                           description="Synthetic mini workload."),
}


def experiment_config(w: Workload, llm_path: str = "<injected>") -> dict:
    """The keys MedTsLLM reads, values of the shipped configs (configs/datasets/*.toml)."""
    return {
        "task": w.task, "model": "medtsllm", "history_len": w.T, "pred_len": w.pred,
        "training": {"dropout": 0.0, "batch_size": w.B, "learning_rate": 1e-4},
        "setup": {"dtype": "mixed", "seed": 0},
        "tasks": {"segmentation": {"mode": "boundary-prediction"}},
        "models": {"medtsllm": {
            "d_model": 32, "d_ff": w.d_ff, "n_heads": 8, "num_tokens": 1024,
            "covariate_mode": w.covariate_mode, "embedding_downsample_mode": "linear",
            "patching": {"patch_len": 16, "stride": 8},
            "prompting": {"dataset": True, "task": True, "clip": False, "input_stats": False, "examples": False,
                          "input_stats_dim": 0, "input_stats_select": "all"},
            "llm": {"enabled": True, "llm": llm_path, "llm_layers": -1, "load_in_4bit": False, "load_in_8bit": False},
            **({"lora": {"enabled": True, "layers": "auto", "rank": w.lora_rank, "alpha": 16, "init": True,
                         "dropout": 0.0, "rslora": True}} if w.lora_rank else {}),
        }},
    }


class AttrDict(dict):
    """Attribute + `.get` + `in` access, the surface of the reference's dict_to_object (utils.py:19-39)."""

    def __init__(self, d):
        super().__init__({k: AttrDict(v) if isinstance(v, dict) else v for k, v in d.items()})

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class SyntheticDataset:
    def __init__(self, w: Workload):
        self.n_features = w.C
        self.n_classes = w.n_classes
        self.description = w.description
        self.task_description = None


class FixedLengthTokenizer:
    """Deterministic stand-in for the HF tokenizer (none exists offline): maps each distinct prompt part
    to a fixed pseudo-random id sequence such that a sample's parts total exactly `prompt_len` tokens
    (SURVEY.md §8d "fixed-Lp=128 variant, random ids, seed 7") — FLOP counts are then tokenizer-
    independent.  Same call surface as the one MedTsLLM uses: `tok(text).input_ids`, bos/eos/pad."""

    bos_token = "<s>"
    eos_token = "</s>"

    def __init__(self, vocab: int, prompt_len: int, n_parts: int = 4, seed: int = 7):
        self.vocab, self.prompt_len, self.n_parts = vocab, prompt_len, n_parts
        self._g = torch.Generator().manual_seed(seed)
        self._parts: dict[str, list[int]] = {}
        self.pad_token = None
        self.pad_token_id = 2

    def __call__(self, text, **_):
        ids = self._parts.get(text)
        if ids is None:
            k = len(self._parts)
            base = self.prompt_len // self.n_parts
            n = base + (self.prompt_len - base * self.n_parts if k == 0 else 0)
            ids = torch.randint(3, self.vocab, (n,), generator=self._g).tolist()
            self._parts[text] = ids

        class _Enc:
            input_ids = ids
        return _Enc()


def make_inputs(w: Workload, seed: int = 1234, device="cpu", pin: bool = False):
    """x = z*scale_c + offset_c, z~N(0,1), scale~U(0.5,5), offset~U(-10,10)  (SURVEY.md §8d)."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(w.B, w.T, w.C, generator=g)
    scale = torch.rand(w.C, generator=g) * 4.5 + 0.5
    offset = torch.rand(w.C, generator=g) * 20 - 10
    x = (z * scale + offset).contiguous()
    if pin and torch.cuda.is_available():
        x = x.pin_memory()
    if device != "cpu":
        x = x.to(device)
    return {"x_enc": x}


def forward_flops(w: Workload) -> dict:
    """Algorithmic forward FLOPs per batch, SURVEY.md §8d formula; returns the total and the GEMM part
    that runs on the tcgen05 kernel (everything except attention and the 3-tap conv)."""
    s = w.backbone
    Bp, L, N, D, V = w.B, w.seq, w.n_patches, s.hidden, s.vocab
    H, E, Sp = 8, w.d_ff, 1024
    dm = 32 * (w.C if w.covariate_mode == "concat" else 1)
    n_out = w.pred * (w.C if w.task in ("forecasting", "reconstruction", "anomaly_detection", "pretraining")
                      else (w.n_classes if w.n_classes > 2 else 1))
    w_blk = 4 * D * D + (3 if s.kind == "llama" else 2) * D * s.inter
    backbone_gemm = 2.0 * Bp * L * w_blk * s.layers
    attn = 4.0 * Bp * L * L * D * s.layers
    mapping = 2.0 * D * V * Sp
    kv = 4.0 * Sp * D * (H * E)
    qo = 2.0 * Bp * N * (dm * H * E + H * E * D)
    rep = 4.0 * Bp * H * N * Sp * E
    conv = 2.0 * w.B * w.C * N * 3 * 16 * 32
    ds = 2.0 * Bp * N * D * E
    head = 2.0 * Bp * E * N * n_out
    gemm = backbone_gemm + mapping + kv + qo + rep + ds + head
    return {"total": gemm + attn + conv, "gemm": gemm, "backbone_gemm": backbone_gemm, "attention": attn,
            "mapping": mapping}
