"""`MedTsLLM` — drop-in for the reference class of the same name (models/medtsllm.py:24-527).

Same constructor `(config, dataset)`, same `forward(inputs: dict) -> Tensor`, same config keys, same
`state_dict()` keys (adapters only; the frozen LLM and `word_embeddings` are never checkpointed,
models/medtsllm.py:235-246), same `load_pretrained`, `supported_tasks`, `lora_enabled`.  Everything
numeric runs on libmtsb200 kernels through `ops` (C ABI): there is no torch-arithmetic fallback and
a CPU tensor raises `MtsError`.

Host-side logic kept in Python exactly like the reference: prompt text building and tokenisation
(HF tokenizer).  The per-sample/per-part embedding gathers + left padding + concat
(models/medtsllm.py:299-311, 331-337, 349) become ONE batched gather kernel over a host-built
token-id table, with static parts tokenised once and cached.
"""
from __future__ import annotations

import math
import os
import weakref

import torch
import torch.nn as nn

from . import ops
from ._lib import BIAS_M, BIAS_N, EPI_RESID_ADD, MtsError
from .graph import GraphReplay
from .backbone import BackboneSpec, KernelBackbone, spec_from_hf_config

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

_DEFAULT_PROMPTING = {"dataset": True, "clip": True, "input_stats": True, "task": True, "examples": False,
                      "input_stats_dim": 0, "input_stats_select": "all"}


def _get(obj, key, default=None):
    """config objects are the reference's `dict_to_object` (utils.py:19-39) or plain dicts."""
    if obj is None:
        return default
    if isinstance(obj, dict):
        return obj.get(key, default)
    if hasattr(obj, "get"):
        return obj.get(key, default)
    return getattr(obj, key, default)


def _has(obj, key):
    if isinstance(obj, dict):
        return key in obj
    try:
        return key in obj
    except TypeError:
        return hasattr(obj, key)


class _TokenEmbeddingParams(nn.Module):
    """Parameter holder with the reference's module path (models/layers/embed.py:29-42):
    `value_embedding.tokenConv.weight` [d_model, patch_len, 3], kaiming-normal init."""

    def __init__(self, c_in, d_model):
        super().__init__()
        self.tokenConv = nn.Conv1d(c_in, d_model, kernel_size=3, padding=1, padding_mode="circular", bias=False)
        nn.init.kaiming_normal_(self.tokenConv.weight, mode="fan_in", nonlinearity="leaky_relu")


class _PatchEmbeddingParams(nn.Module):
    def __init__(self, d_model, patch_len, stride, dropout):
        super().__init__()
        self.patch_len, self.stride, self.p_dropout = patch_len, stride, dropout
        self.value_embedding = _TokenEmbeddingParams(patch_len, d_model)


class _ReprogrammingParams(nn.Module):
    """models/medtsllm.py:555-564 — parameter holder (same names, same nn.Linear init)."""

    def __init__(self, d_model, n_heads, d_keys, d_llm):
        super().__init__()
        self.query_projection = nn.Linear(d_model, d_keys * n_heads)
        self.key_projection = nn.Linear(d_llm, d_keys * n_heads)
        self.value_projection = nn.Linear(d_llm, d_keys * n_heads)
        self.out_projection = nn.Linear(d_keys * n_heads, d_llm)
        self.n_heads = n_heads


class _FlattenHeadParams(nn.Module):
    """models/medtsllm.py:541-546."""

    def __init__(self, nf, target_window):
        super().__init__()
        self.linear = nn.Linear(nf, target_window)


class _LLMHandle:
    """What the reference's surroundings touch on `model.llm` (loggers/base_logger.py:42-43,
    models/medtsllm.py:346): `.config` and, for LoRA runs, `.save_pretrained`."""

    def __init__(self, config):
        self.config = config

    def save_pretrained(self, path, *a, **k):
        raise MtsError("model.llm.save_pretrained is only meaningful with lora.enabled (loggers/base_logger.py:42-43)")


_LIVE_MODELS: "weakref.WeakSet" = weakref.WeakSet()
_HOOK_INSTALLED = False


def _install_optimizer_hook():
    """Every `optimizer.step()` of any torch optimizer bumps `_opt_steps` of the live models, which is part of every
    weight-cache key: optimizers that update through `p.data` (no `_version` bump — older / third-party ones such as
    the Ranger21 the reference allows, tasks/base.py:12) still invalidate the bf16 copies and captured graphs."""
    global _HOOK_INSTALLED
    if _HOOK_INSTALLED:
        return
    try:
        from torch.optim.optimizer import register_optimizer_step_post_hook
    except ImportError:      # very old torch: `model.invalidate_caches()` stays available
        return

    def _bump(optimizer, args, kwargs):
        for m in list(_LIVE_MODELS):
            m._opt_steps += 1
            if getattr(m, "lora_enabled", False):
                m.llm._opt_steps += 1

    register_optimizer_step_post_hook(_bump)
    _HOOK_INSTALLED = True


class MedTsLLM(nn.Module):

    supported_tasks = ["forecasting", "reconstruction", "anomaly_detection", "semantic_segmentation", "segmentation", "pretraining"]
    supported_modes = ["univariate", "multivariate"]

    def __init__(self, config, dataset, backbone: KernelBackbone | None = None, tokenizer=None):
        """`backbone`/`tokenizer` are optional injection points for benchmarks and tests (a
        device-resident random-init stack); by default both are loaded exactly as the reference
        does (AutoConfig/AutoModel/AutoTokenizer from `llm.llm`, models/medtsllm.py:129-233)."""
        super().__init__()
        self.config = config
        models_cfg = _get(config, "models")
        self.model_config = _get(models_cfg, "medtsllm") if _has(models_cfg, "medtsllm") else _get(models_cfg, "timellm")
        mc = self.model_config

        self.device = None
        self.pred_len = _get(config, "pred_len")
        self.seq_len = _get(config, "history_len")
        self.task = _get(config, "task")
        self.task_description = self.get_task_description(dataset)
        self.dataset_description = dataset.description

        self.d_ff = _get(mc, "d_ff")
        self.d_model = _get(mc, "d_model")
        self.n_attention_heads = _get(mc, "n_heads")
        self.num_tokens = _get(mc, "num_tokens")
        self.dropout = _get(_get(config, "training"), "dropout")
        self.n_lags = 5

        patching = _get(mc, "patching")
        self.patch_len = _get(patching, "patch_len")
        self.stride = _get(patching, "stride")
        self.n_patches = int((self.seq_len - self.patch_len) / self.stride + 2)
        self.d_patch = self.d_model

        self.covariate_mode = _get(mc, "covariate_mode")
        self.n_features = dataset.n_features
        self.n_classes = dataset.n_classes if self.task in ["classification", "semantic_segmentation"] else 0

        if self.task in ["forecasting", "reconstruction", "anomaly_detection", "pretraining"]:
            self.n_outputs_per_step = self.n_features
        elif self.task == "semantic_segmentation":
            self.n_outputs_per_step = self.n_classes if self.n_classes > 2 else 1
        elif self.task == "segmentation":
            self.n_outputs_per_step = 1
            self.seg_mode = _get(_get(_get(config, "tasks"), "segmentation"), "mode")
            assert self.seg_mode in ["boundary-prediction", "steps-to-boundary"]
        else:
            raise ValueError(f"Task {self.task} is not supported.")
        self.n_outputs = self.n_outputs_per_step * self.pred_len

        # covariate modes (models/medtsllm.py:71-87)
        if self.covariate_mode == "univariate":
            assert self.n_features == 1
        elif self.covariate_mode == "interleave":
            self.n_patches *= self.n_features
        elif self.covariate_mode == "concat":
            self.d_model *= self.n_features
        elif self.covariate_mode == "independent" or self.covariate_mode == "add":
            pass
        elif self.covariate_mode == "merge-end":
            self.feature_weighting = nn.Linear(self.n_features * self.n_outputs_per_step, self.n_outputs_per_step)
        elif self.covariate_mode == "weighted-average":
            self.feature_weighting = nn.Linear(self.n_features, 1)
        else:
            raise ValueError(f"Unknown covariate mode {self.covariate_mode}")

        self.setup_llm(backbone, tokenizer)

        # trainable adapters: same module paths / shapes / init as the reference (models/medtsllm.py:91-101)
        self.mapping_layer = nn.Linear(self.vocab_size, self.num_tokens)
        self.patch_embedding = _PatchEmbeddingParams(self.d_patch, self.patch_len, self.stride, self.dropout)
        self.reprogramming_layer = _ReprogrammingParams(self.d_model, self.n_attention_heads, self.d_ff, self.d_llm)
        self.output_projection = _FlattenHeadParams(self.d_ff * self.n_patches, self.n_outputs)

        self.embedding_downsample_mode = _get(mc, "embedding_downsample_mode")
        if self.embedding_downsample_mode == "linear":
            self.embedding_downsample_layer = nn.Linear(self.d_llm, self.d_ff)
        elif self.embedding_downsample_mode == "average":
            assert self.d_llm % self.d_ff == 0                                 # models/medtsllm.py:100-101
        elif self.embedding_downsample_mode != "truncate":
            raise ValueError(f"Unknown embedding downsample mode {self.embedding_downsample_mode}")
        self._ds_fixed = None      # constant [d_ff, D] selection / averaging matrix for truncate / average

        if self.dropout and self.dropout > 0:
            # PatchEmbedding / reprogramming dropout (models/medtsllm.py:93-94) are training-time noise;
            # the kernels implement the deterministic path.  Recorded so train() can refuse loudly.
            self._dropout_requested = float(self.dropout)
        else:
            self._dropout_requested = 0.0

        # in-batch prompt de-duplication (shared-prefix row layout, include/mts_b200.h): on by default,
        # MTS_SHARE_PREFIX=0 or `model.share_prompt_prefix = False` computes every sample's prompt rows
        self.share_prompt_prefix = os.environ.get("MTS_SHARE_PREFIX", "1") != "0"
        # inference replays a captured CUDA graph of the whole path once the same (shape, prompt table, weights)
        # has been seen twice in a row (MTS_CUDA_GRAPH=0 / `model.use_cuda_graph = False`: launch kernel by kernel)
        self.use_cuda_graph = os.environ.get("MTS_CUDA_GRAPH", "1") != "0"
        self._graph = GraphReplay()
        # training steps replay two captured graphs (forward with its stash, backward chain) where the step is bound by
        # the host issuing launches: "auto" = single process, or a GPT-2-sized backbone / LoRA under data parallelism
        # (the gradient all-reduce then runs after the backward graph instead of underneath it); "1" / "0" force it
        self.use_train_graph = os.environ.get("MTS_TRAIN_GRAPH", "auto")
        self._train_graph = None
        self._force_recast = False
        self._capture = None       # tests: dict filled with per-stage tensors
        self._ids_cache = None     # (prompt parts, host id table, device id table)
        self._prompt_cache: dict[str, list[int]] = {}
        self._src_cache = None     # (versions, source_bf16, K_bf16, Vt_bf16)
        self._w_cache: dict[str, tuple[int, torch.Tensor]] = {}
        # generation of the device-side caches (bf16 weight copies, prototype K/V, constant down-sample matrix, RoPE
        # tables, LoRA operands): part of the graph-replay key, because a captured graph holds raw pointers into them
        self._cache_gen = 0
        self._opt_steps = 0        # optimizer steps seen by the global hook (part of every weight-cache key)
        _LIVE_MODELS.add(self)
        _install_optimizer_hook()

    # ------------------------------------------------------------------------------------------ LLM
    def setup_llm(self, backbone=None, tokenizer=None):
        mc = self.model_config
        llm_cfg = _get(mc, "llm")
        self.llm_enabled = _get(llm_cfg, "enabled")
        if not self.llm_enabled:
            # (the reference builds an `llm_replacement` MLP for this flag, models/medtsllm.py:103-109, but never calls
            # it: what the flag really does is skip the freeze at :227-233, i.e. fine-tune the whole LLM)
            raise NotImplementedError("llm.enabled = false leaves the whole backbone trainable (models/medtsllm.py:227-233); "
                                      "backbone weight gradients are outside the accelerated path")
        self.llm_id = _get(llm_cfg, "llm")
        self.llm_layers = _get(llm_cfg, "llm_layers")
        if _get(llm_cfg, "load_in_4bit") or _get(llm_cfg, "load_in_8bit"):
            raise NotImplementedError("bitsandbytes 4/8-bit loading is outside the accelerated path")
        lora = _get(mc, "lora")
        self.lora_enabled = bool(lora is not None and _get(lora, "enabled"))
        if self.lora_enabled:
            assert _get(lora, "layers") == "auto"                         # models/medtsllm.py:190
            if _get(lora, "dropout", 0.0):
                raise NotImplementedError("lora.dropout > 0 (dropout is not implemented on the kernel path)")
        dtype_name = _get(_get(self.config, "setup"), "dtype")
        if dtype_name not in ("float32", "float", "fp32", "32", 32, "mixed"):
            raise NotImplementedError(
                f"setup.dtype={dtype_name!r}: adapters are fp32 masters and kernels compute in bf16 with fp32 "
                "accumulation; use 'mixed' (the shipped configs) or 'float32'")

        # precision of evaluation-mode forwards (precise.py): "bf16" = the fast path (the reference's bf16-autocast
        # numerics); "tf32" = the reference's own evaluation regime (fp32 weights, TF32 matmuls, tasks/base.py:19-22);
        # "fp32" = fp32-grade contractions (3xTF32).  `setup.dtype = "float32"` asks the reference for fp32 weights
        # without autocast, so it maps to "tf32"; "mixed" (the shipped configs) stays on the fast path.  MTS_PRECISION
        # overrides; an injected backbone decides by how it was built.  Training always runs the bf16 path.
        want = os.environ.get("MTS_PRECISION") or ("bf16" if dtype_name == "mixed" else "tf32")
        if want not in ("bf16", "tf32", "fp32"):
            raise ValueError(f"MTS_PRECISION={want!r}: expected bf16, tf32 or fp32")
        self.precision = want if backbone is None else backbone.precision
        if backbone is not None:
            self._backbone = backbone
            object.__setattr__(self, "_hf_model", None)
            self.tokenizer = tokenizer
            self.llm = _LLMHandle(None)
            spec = backbone.spec
        else:
            from transformers import AutoConfig, AutoModel, AutoTokenizer
            cache_dir = _get(_get(self.config, "paths", {}), "llm_path")
            if cache_dir in ("", "none"):
                cache_dir = None
            llm_config = AutoConfig.from_pretrained(self.llm_id, cache_dir=cache_dir)
            if self.llm_layers > 0 and self.llm_layers < llm_config.num_hidden_layers:
                llm_config.num_hidden_layers = self.llm_layers
            spec = spec_from_hf_config(llm_config)
            # host copy, converted to kernel layout when the model is moved to the GPU (`_apply`)
            # (kept out of the Module tree: the frozen LLM is never a parameter of this model)
            object.__setattr__(self, "_hf_model", AutoModel.from_pretrained(
                self.llm_id, config=llm_config, torch_dtype=torch.float32, cache_dir=cache_dir))
            self._backbone = None
            self.tokenizer = tokenizer or AutoTokenizer.from_pretrained(self.llm_id, cache_dir=cache_dir)
            self.llm = _LLMHandle(llm_config)

        if self.tokenizer is not None:
            if self.tokenizer.eos_token:
                self.tokenizer.pad_token = self.tokenizer.eos_token
            else:
                self.tokenizer.add_special_tokens({"pad_token": "[PAD]"})
                self.tokenizer.pad_token = "[PAD]"
        if spec.vocab > 100_000:
            raise NotImplementedError("vocabularies > 100k (sub-sampled trainable word_embeddings, "
                                      "models/medtsllm.py:219-222) are outside the BASELINE configs")
        if self.lora_enabled:
            from .lora import LoraAdapters
            if spec.kv_heads:
                raise NotImplementedError("LoRA on a grouped-query backbone (v_proj is narrower than the kernel's expanded "
                                          "K / V projection) is outside the BASELINE configs")
            init = _get(lora, "init", True)
            if init not in (True, False):
                raise NotImplementedError(f"lora.init = {init!r}")
            llm_config = self.llm.config
            self.llm = LoraAdapters(spec, _get(lora, "rank"), _get(lora, "alpha"), rslora=_get(lora, "rslora", True),
                                    init=bool(init))
            self.llm.config = llm_config
        self.backbone_spec: BackboneSpec = spec
        self.vocab_size = spec.vocab
        self.d_llm = spec.hidden

    def _apply(self, fn, *args, **kwargs):
        """`.to(device, dtype)` from the Trainer (tasks/base.py:41): adapters follow torch; the frozen
        stack is converted once to the bf16 kernel layout on the target GPU and the host copy freed."""
        super()._apply(fn, *args, **kwargs)
        dev = self.mapping_layer.weight.device
        if dev.type == "cuda":
            if self._backbone is None:
                self._backbone = KernelBackbone.from_hf(self._hf_model, dev, precision=self.precision)
                object.__setattr__(self, "_hf_model", None)
            elif self._backbone.device != dev:
                raise MtsError("the kernel backbone lives on another device")
            self.device = dev
        return self

    def state_dict(self, *args, **kwargs):
        # adapters only, as the reference (models/medtsllm.py:235-246): the frozen backbone is not a Module and
        # `llm.*` (LoRA pairs; saved separately through llm.save_pretrained, loggers/base_logger.py:42-43) is dropped
        sd = super().state_dict(*args, **kwargs)
        for k in [k for k in sd if k.startswith("llm.")]:
            del sd[k]
        return sd

    def load_pretrained(self, saved_state):
        """models/medtsllm.py:515-527."""
        for k in ("word_embeddings", "output_projection.linear.bias", "output_projection.linear.weight"):
            saved_state.pop(k, None)
        incompat = self.load_state_dict(saved_state, strict=False)
        assert len(incompat.unexpected_keys) == 0, f"Unexpected keys in model state: {incompat.unexpected_keys}"
        return list(saved_state.keys())

    def train(self, mode: bool = True):
        return super().train(mode)

    # ------------------------------------------------------------------------------------------ prompt
    def get_task_description(self, dataset):
        """models/medtsllm.py:497-513."""
        if getattr(dataset, "task_description", None) is not None:
            return dataset.task_description
        if self.task in ("forecasting", "pretraining"):
            return f"Forecast the next {self.pred_len} steps given the previous {self.seq_len} steps of data."
        if self.task in ("anomaly_detection", "reconstruction"):
            return f"Reconstruct the past {self.seq_len} steps of data as accurately as possible using the following information."
        if self.task == "semantic_segmentation":
            return f"Classify the past {self.seq_len} steps of data as accurately as possible using the following information."
        if self.task == "segmentation":
            return f"Identify the change points in the past {self.seq_len} steps of data to segment the sequence."
        raise ValueError(f"Task {self.task} is not supported.")

    def build_prompt(self, inputs):
        """models/medtsllm.py:386-439 — per-sample list of prompt parts (host strings)."""
        bs = inputs["x_enc"].size(0)
        cfg = _get(self.model_config, "prompting") or _DEFAULT_PROMPTING
        flags = {k: _get(cfg, k, _DEFAULT_PROMPTING[k]) for k in _DEFAULT_PROMPTING}
        if not (flags["dataset"] or flags["clip"] or flags["input_stats"] or flags["task"] or flags["examples"]):
            return [[] for _ in range(bs)]
        dataset_prompt = f"Dataset: {self.dataset_description}" if flags["dataset"] else ""
        # prompting.examples (:402-405, datasets/ecg.py:140-166): per sample a tuple of parts, strings or time-series
        # tensors [1, T_ex, C] that `encode_part` (:313-319) runs through encode_ts like the window itself
        example_prompts = inputs["examples"] if flags["examples"] else [("",)] * bs
        clip_prompts = inputs.get("descriptions", [""] * bs) if flags["clip"] else [""] * bs
        stats_prompts = self.build_input_stats_prompt(flags, inputs) if flags["input_stats"] else [""] * bs
        task_prompt = f"Task: {self.task_description}" if flags["task"] else ""
        bos = self.tokenizer.bos_token if self.tokenizer.bos_token is not None else ""
        prompts = []
        for b in range(bs):
            parts = [bos, dataset_prompt, *example_prompts[b], clip_prompts[b], stats_prompts[b], task_prompt, "Time series:"]
            parts = [p for p in parts if not (isinstance(p, str) and p == "")]
            parts = [(p + " " if isinstance(p, str) and i != 0 else p) for i, p in enumerate(parts)]
            prompts.append(parts)
        return prompts

    def build_input_stats_prompt(self, flags, inputs):
        """models/medtsllm.py:441-495 — text statistics of the window.  Text generation on the host
        side of the boundary; the numbers come from the same torch reductions as the reference."""
        xs = inputs["x_enc"].detach()
        if xs.ndim == 2:
            xs = xs.unsqueeze(-1)
        assert flags["input_stats_select"] == "all"

        def fmt_list(v):
            return "[" + ", ".join(v) + "]"

        def fmt_float(v):
            return fmt_list([fmt_float(u) for u in v]) if isinstance(v, list) else f"{v:.3f}"

        def fmt_trend(v):
            if v is True:
                return "upward"
            if v is False:
                return "downward"
            if isinstance(v, list):
                return fmt_list([fmt_trend(u) for u in v])
            return v

        if flags["input_stats_dim"] == "all":
            insert, s = "per feature", "s"
        else:
            d = flags["input_stats_dim"]
            insert, s = f"feature {d}", ""
            xs = xs[:, :, d]
        B, T = xs.size(0), xs.size(1)
        with torch.no_grad():
            if xs.is_cuda and xs.dtype == torch.float32 and T % 2 == 0 and T <= 12288:
                # one fused kernel pair + ONE packed read-back (the reference: five `.tolist()` device syncs, :477-481)
                x3 = inputs["x_enc"].detach()
                x3 = x3.unsqueeze(-1) if x3.ndim == 2 else x3
                if flags["input_stats_dim"] == "all":
                    stats, lags_d = ops.input_stats(x3, n_lags=self.n_lags)
                else:
                    stats, lags_d = ops.input_stats(x3, f0=int(flags["input_stats_dim"]), n_features=1, n_lags=self.n_lags)
                w = stats.shape[1]
                packed = torch.cat([stats.reshape(B, 4 * w), lags_d.to(torch.float32)], dim=1).cpu()
                st = packed[:, :4 * w].reshape(B, w, 4)
                pick = (lambda k: st[:, :, k]) if flags["input_stats_dim"] == "all" else (lambda k: st[:, 0, k])
                mins, maxs, meds = (pick(k).tolist() for k in range(3))
                trends = (pick(3) > 0.5).tolist()
                lags = packed[:, 4 * w:].long().tolist()
            else:
                # odd window lengths (irfft then returns T-1 points) and host tensors: the reference's own torch route,
                # with its five read-backs packed into one
                mins, maxs = torch.min(xs, dim=1).values, torch.max(xs, dim=1).values
                meds = torch.median(xs.float(), dim=1).values
                trends = xs.diff(dim=1).sum(dim=1) > 0
                lags = _calcute_lags(xs.float(), self.n_lags)                 # [B, n_lags]
                packed = torch.cat([t.reshape(B, -1).double() for t in (mins, maxs, meds, trends, lags)], dim=1).cpu()
                w = mins.reshape(B, -1).shape[1]
                shape = list(mins.shape[1:])
                unpack = lambda k: packed[:, k * w:(k + 1) * w].reshape([B] + shape)      # noqa: E731
                mins, maxs, meds = (unpack(k).to(xs.dtype).tolist() for k in range(3))
                trends = unpack(3).bool().tolist()
                lags = packed[:, 4 * w:].long().tolist()
        return [
            f"Input statistics ({insert}): min value{s} = {fmt_float(mins[b])}, max value{s} = {fmt_float(maxs[b])}, "
            f"median value{s} = {fmt_float(meds[b])}, the trend of input is {fmt_trend(trends[b])}, "
            f"the top {self.n_lags} lags are {lags[b]}."
            for b in range(xs.size(0))
        ]

    def _tokenize_part(self, text: str) -> list[int]:
        ids = self._prompt_cache.get(text)
        if ids is None:
            # same call as encode_text (models/medtsllm.py:300): default add_special_tokens
            ids = list(self.tokenizer(text, padding=False, truncation=False).input_ids)
            if len(self._prompt_cache) < 4096:
                self._prompt_cache[text] = ids
        return ids

    def prompt_token_ids(self, inputs):
        """Host token-id table [B, Lp] (int32, LEFT-padded with the pad id, models/medtsllm.py:304-311).  Positions that
        hold a time-series example part instead of a token carry a negative id (-(b+1): distinct per sample, so they
        never count as a shared prefix); the parts themselves hang off the table as `.example_segments` =
        [(sample, first position, tensor [1, T_ex, C])]."""
        prompts = self.build_prompt(inputs)
        has_ts = any(isinstance(p, torch.Tensor) for parts in prompts for p in parts)
        key = None if has_ts else tuple(tuple(parts) for parts in prompts)
        if key is not None and self._ids_cache is not None and self._ids_cache[0] == key:   # static prompts: same table every batch
            return self._ids_cache[1]
        per_sample, segs = [], []
        for b, parts in enumerate(prompts):
            row = []
            for part in parts:
                if isinstance(part, torch.Tensor):
                    if self.covariate_mode not in ("concat", "univariate"):
                        raise NotImplementedError("prompting.examples needs covariate_mode concat / univariate (the "
                                                  "reference's torch.cat of prompt parts breaks for the others, :332)")
                    ts = part.unsqueeze(-1) if part.ndim == 2 else part
                    if ts.ndim != 3 or ts.shape[0] != 1 or ts.shape[2] != self.n_features:
                        raise ValueError(f"example part must be [1, T, {self.n_features}], got {tuple(part.shape)}")
                    n = ops.n_patches(ts.shape[1], self.patch_len, self.stride)
                    segs.append((b, len(row), ts))
                    row.extend([-(b + 1)] * n)
                else:
                    row.extend(self._tokenize_part(part))
            per_sample.append(row)
        Lp = max((len(p) for p in per_sample), default=0)
        pad = self.tokenizer.pad_token_id if self.tokenizer is not None else 0
        table = torch.full((len(per_sample), Lp), pad if pad is not None else 0, dtype=torch.int32)
        for b, ids in enumerate(per_sample):
            if ids:
                table[b, Lp - len(ids):] = torch.tensor(ids, dtype=torch.int32)
        table.example_segments = [(b, Lp - len(per_sample[b]) + off, ts) for b, off, ts in segs]
        self._ids_cache = (key if key is not None else object(), table, None)   # (+ shared-prefix length once measured)
        return table

    def _ids_device(self, ids: torch.Tensor, dev) -> torch.Tensor:
        """Device copy of the host id table, re-used while the prompts do not change (dataset / task prompts are static;
        only clip descriptions, input statistics and examples vary per batch) — also what keeps host-to-device copies
        out of a CUDA-graph capture."""
        c = self._ids_cache
        if c is not None and c[1] is ids and c[2] is not None and c[2].device == dev:
            return c[2]
        ids_dev = ids.to(dev, non_blocking=True)
        if c is not None and c[1] is ids:
            self._ids_cache = (c[0], c[1], ids_dev) + tuple(c[3:])
        return ids_dev

    def _shared_prefix_len(self, ids: torch.Tensor, Bp: int, L: int, precise: bool = False) -> int:
        """Number of leading prompt positions (left padding included) that hold the same token in every
        sample of the batch.  The backbone is causal and the reference passes no padding mask
        (models/medtsllm.py:350), so the hidden states of those positions are the same for every sample
        at every layer: they are computed once per batch instead of once per sample.  0 = plain layout."""
        if not self.share_prompt_prefix or Bp < 2 or ids.shape[1] == 0:
            return 0
        c = self._ids_cache
        if c is not None and c[1] is ids and len(c) > 3:
            Lc = c[3]
        else:
            same = (ids == ids[0:1]).all(dim=0)
            Lc = int(same.to(torch.int8).cumprod(0).sum())
            if c is not None and c[1] is ids:
                self._ids_cache = (c[0], c[1], c[2], Lc)
        if Lc < 16:
            return 0
        if precise:           # the fp32 attention kernel tiles the keys: no length limit
            return Lc
        if 240 < L <= 256:
            # the tensor-memory attention kernel takes sequences of 241..256 positions only with a prefix that is a
            # multiple of 16 rows (csrc/attention_tc.cu: attn_tc_eligible); the per-sample layout of the same model always
            # takes it there, so give the shared layout the same route (identical outputs): a few prefix rows become own rows
            Lc -= Lc % 16
        # the sequence-resident attention kernels keep all L positions of one head in shared memory
        hd = self.backbone_spec.head_dim
        L64 = (L + 63) // 64 * 64
        smem = (2 * L64 + 256) * (hd + 8) * 2 + 2 * L64 * 4 + 16
        return Lc if smem <= 220 * 1024 else 0

    @staticmethod
    def _expand_rows(t: torch.Tensor, Bp: int, L: int, Lc: int) -> torch.Tensor:
        """[Lc + Bp*(L-Lc), D] shared-prefix rows -> [Bp, L, D] (tests / captures only)."""
        D = t.shape[-1]
        if Lc == 0:
            return t.view(Bp, L, D).clone()
        return torch.cat([t[:Lc].unsqueeze(0).expand(Bp, Lc, D), t[Lc:].view(Bp, L - Lc, D)], dim=1).contiguous()

    # ------------------------------------------------------------------------------------------ weights
    PARAM_ORDER = (
        "patch_embedding.value_embedding.tokenConv.weight",
        "mapping_layer.weight", "mapping_layer.bias",
        "reprogramming_layer.query_projection.weight", "reprogramming_layer.query_projection.bias",
        "reprogramming_layer.key_projection.weight", "reprogramming_layer.key_projection.bias",
        "reprogramming_layer.value_projection.weight", "reprogramming_layer.value_projection.bias",
        "reprogramming_layer.out_projection.weight", "reprogramming_layer.out_projection.bias",
        "embedding_downsample_layer.weight", "embedding_downsample_layer.bias",
        "output_projection.linear.weight", "output_projection.linear.bias",
        "feature_weighting.weight", "feature_weighting.bias",
    )

    def param_order(self):
        """PARAM_ORDER restricted to the tensors this configuration has (no down-sample Linear for the
        truncate / average modes)."""
        named = dict(self.named_parameters())
        return [k for k in self.PARAM_ORDER if k in named]

    def adapter_params(self):
        """The trainable tensors (15 adapter tensors with the shipped configs) + LoRA pairs if enabled."""
        named = dict(self.named_parameters())
        return [named[k] for k in self.param_order()] + (self.llm.params() if self.lora_enabled else [])

    def _downsample_operands(self):
        """(W bf16 [d_ff, ld], bias fp32 or None) of the down-sample step (models/medtsllm.py:354-363).
        `linear` is the trainable Linear; `truncate` (dec[:, :, :d_ff]) and `average` (mean over groups of
        D/d_ff consecutive features) are the same GEMM with a constant selection / averaging matrix — exact
        in bf16 (entries 1 or 1/g, fp32 accumulation)."""
        if self.embedding_downsample_mode == "linear":
            return (self._bf16_weight("wds", self.embedding_downsample_layer.weight),
                    self.embedding_downsample_layer.bias.detach())
        if self._ds_fixed is None or self._ds_fixed.device != self.device:
            E, D = self.d_ff, self.d_llm
            w = torch.zeros(E, D, dtype=torch.float32)
            if self.embedding_downsample_mode == "truncate":
                w[torch.arange(E), torch.arange(E)] = 1.0
            else:
                g = D // E
                w.view(E, E, g)[torch.arange(E), torch.arange(E), :] = 1.0 / g
            self._ds_fixed = ops.cast_bf16(w.to(self.device))
            self._cache_gen += 1
        return self._ds_fixed, None

    def train_graph_enabled(self) -> bool:
        """Whether this training forward may run on the captured step graphs (train.TrainGraph)."""
        mode = self.use_train_graph
        if (mode in (False, "0", "off") or self._capture is not None or self._dropout_requested > 0
                or self._has_backbone_dropout()):
            return False            # (dropout draws fresh host seeds every step)
        if mode == "auto":
            from . import dp
            if dp.is_active() and not (self.lora_enabled or self.backbone_spec.hidden < 2048):
                return False        # GPU-bound step: keep the all-reduce overlapped with the backbone dgrad
        if self._train_graph is None:
            from .train import TrainGraph
            self._train_graph = TrainGraph()
        return True

    def _set_force_recast(self, flag: bool):
        self._force_recast = flag
        if self.lora_enabled:
            self.llm._force_recast = flag

    def _bf16_weight(self, name: str, p: torch.Tensor) -> torch.Tensor:
        """bf16 copy [rows, ceil8(cols)] (zero padded: TMA rows must be 16-byte aligned) of a trainable
        fp32 master, re-cast by our kernel only when the optimizer has changed it (`_version` bumps on
        every in-place update)."""
        key = (p._version, p.data_ptr(), self._opt_steps)
        hit = self._w_cache.get(name)
        if hit is not None and hit[0] == key and not self._force_recast:
            return hit[1]
        rows, cols = p.shape
        w = ops.cast_rows(p.detach().contiguous(), rows=rows, cols=cols)
        self._w_cache[name] = (key, w)
        self._cache_gen += 1
        return w

    def set_precision(self, precision: str):
        """Switch the numerics of evaluation-mode forwards ("bf16" | "tf32" | "fp32", see precise.py).  The parity modes
        need a backbone built with fp32 operands (MTS_PRECISION / setup.dtype at load time, or an injected stack)."""
        if precision not in ("bf16", "tf32", "fp32"):
            raise ValueError(f"precision {precision!r}: expected bf16, tf32 or fp32")
        if precision != "bf16":
            if self._backbone is None:
                raise MtsError("move the model to the GPU first")
            self._backbone.set_precision(precision)
        self.precision = precision
        self._cache_gen += 1

    def invalidate_caches(self):
        """Drops every derived device-side copy (bf16 adapter weights, prototype K / V, LoRA operands) and the captured
        graphs.  Needed only after weight surgery that neither bumps `Parameter._version` nor goes through a torch
        optimizer's `step()` (e.g. `p.data.copy_(...)` by hand, EMA weight swaps): see INTEGRATION.md."""
        self._w_cache.clear()
        self._src_cache = None
        self._cache_gen += 1
        self._graph = GraphReplay()
        self._train_graph = None
        if self.lora_enabled:
            self.llm._cache.clear()
            self.llm._fold_cache.clear()

    def _source_kv(self):
        """Prototype path (models/medtsllm.py:281, :574-575): source = W_map E + b; K = W_k source + b_k;
        V^T = W_v source^T + b_v.  Batch independent -> cached until a contributing weight changes."""
        rl = self.reprogramming_layer
        plist = [self.mapping_layer.weight, self.mapping_layer.bias, rl.key_projection.weight,
                 rl.key_projection.bias, rl.value_projection.weight, rl.value_projection.bias]
        key = tuple((p._version, p.data_ptr()) for p in plist) + (self._opt_steps,)
        if self._src_cache is not None and self._src_cache[0] == key and not self._force_recast:
            return self._src_cache[1:]
        self._cache_gen += 1
        bb = self._backbone
        S, D, HE, V = self.num_tokens, self.d_llm, self.d_ff * self.n_attention_heads, self.vocab_size
        dev = self.device
        w_map = self._bf16_weight("map", self.mapping_layer.weight)                       # [S, ceil8(V)]
        source = torch.empty(S, D, device=dev, dtype=torch.bfloat16)
        ops.gemm(w_map, bb.embed_t, source, m=S, n=D, k=V, lda=w_map.shape[1], ldb=bb.embed_t.shape[1],
                 bias=self.mapping_layer.bias.detach(), bias_axis=BIAS_M)
        wk = self._bf16_weight("wk", rl.key_projection.weight)                            # [HE, D]
        wv = self._bf16_weight("wv", rl.value_projection.weight)
        K = torch.empty(S, HE, device=dev, dtype=torch.bfloat16)
        ops.gemm(source, wk, K, m=S, n=HE, k=D, bias=rl.key_projection.bias.detach(), bias_axis=BIAS_N)
        Vt = torch.empty(HE, S, device=dev, dtype=torch.bfloat16)                          # [H*E, S]
        ops.gemm(wv, source, Vt, m=HE, n=S, k=D, bias=rl.value_projection.bias.detach(), bias_axis=BIAS_M)
        self._src_cache = (key, source, K, Vt)
        return source, K, Vt

    # ------------------------------------------------------------------------------------------ forward
    accepts_host_mirror = True      # plugin.HostMirrorBatch (evaluation batches whose host copies stay addressable)

    def forward(self, inputs):
        dev_batch = getattr(inputs, "device_batch", None)
        if dev_batch is not None:
            # plugin.HostMirrorBatch: compute on the device copies; in evaluation hand the predictions back on the host
            # with ONE copy, so that the per-sample `.cpu()` of the reference's predict() loops are no-ops
            if self.training:
                return self.forward(dev_batch)
            out = self.predict(dev_batch)
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(out.device).synchronize()
            return host
        # autograd path (adapter gradients) for training-mode forwards under enabled gradients — what the reference's
        # Trainers differentiate (tasks/forecasting.py:18-26).  Evaluation-mode forwards take the inference path, which
        # also applies the eval-only sigmoid / softmax (models/medtsllm.py:251-259) and returns a tensor without a graph.
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import forward_train
            return forward_train(self, inputs)
        return self.predict(inputs)

    @torch.no_grad()
    def predict(self, inputs):
        """Inference path: models/medtsllm.py:321-384 + the eval-only activation of :248-261.

        The ~8 launches per backbone layer are short next to their launch cost for the GPT-2 configs and for
        batches whose prompt rows are shared, so a steady stream of same-shaped batches (the Trainer's val / test
        loops, tasks/forecasting.py:55-78) replays ONE captured CUDA graph: the windows are copied into the
        graph's input buffer, the graph runs, the predictions are copied out."""
        if (not self.use_cuda_graph or self._capture is not None
                or (self.training and (self._dropout_requested > 0 or self._has_backbone_dropout()))):
            return self._predict_eager(inputs)          # (train-mode dropout draws fresh host seeds every call)
        x_enc = self._check_input(inputs)
        ids = self.prompt_token_ids(inputs)
        params = list(self.parameters()) + (self.llm.params() if self.lora_enabled else [])
        key = (tuple(x_enc.shape), x_enc.device.index, self.training, id(ids), self.share_prompt_prefix, self.precision,
               tuple((p._version, p.data_ptr()) for p in params), self._opt_steps, self._cache_gen,
               self._backbone.cache_gen, self.llm._cache_gen if self.lora_enabled else 0)
        return self._graph.run(key, x_enc, lambda xs: self._predict_eager({**inputs, "x_enc": xs}, ids), hold=ids)

    def _predict_eager(self, inputs, ids=None):
        if self.precision != "bf16" and not self.training:
            from .precise import forward_precise
            out = forward_precise(self, inputs, ids)
        else:
            out = self._forward_impl(inputs, None, ids)
        if not self.training:
            if self.task == "semantic_segmentation":
                if self.n_classes > 2:
                    ops.softmax_lastdim_(out)
                else:
                    ops.sigmoid_(out)
            elif self.task == "segmentation" and self.seg_mode == "boundary-prediction":
                ops.sigmoid_(out)
        return out

    def _backbone_dropout(self):
        """The frozen backbone's own train-mode dropouts (tasks/forecasting.py:18: `model.train()` flips the HF module
        too, so GPT-2 checkpoints' embd / attn / resid 0.1 are live while the reference trains; Llama-2 carries
        attention_dropout 0).  Returns {"embd", "attn", "resid", "seeds"} with fresh seeds from torch's generator, or None
        in evaluation / when every probability is zero.  `self.backbone_dropout = {...}` overrides the HF config (injected
        backbones carry none)."""
        if not self.training:
            return None
        probs = getattr(self, "backbone_dropout", None)
        if probs is None:
            cfg = getattr(self.llm, "config", None)
            g = lambda k: float(getattr(cfg, k, 0.0) or 0.0) if cfg is not None else 0.0      # noqa: E731
            if self.backbone_spec.kind == "gpt2":
                probs = {"embd": g("embd_pdrop"), "attn": g("attn_pdrop"), "resid": g("resid_pdrop")}
            else:
                probs = {"embd": 0.0, "attn": g("attention_dropout"), "resid": 0.0}
        if max(probs["embd"], probs["attn"], probs["resid"]) <= 0:
            return None
        n = 1 + 3 * self.backbone_spec.layers
        return {**probs, "seeds": [int(v) for v in torch.randint(0, 2 ** 62, (n,))]}

    def _has_backbone_dropout(self) -> bool:
        probs = getattr(self, "backbone_dropout", None)
        if probs is not None:
            return max(probs.values()) > 0
        cfg = getattr(self.llm, "config", None)
        if cfg is None:
            return False
        keys = ("embd_pdrop", "attn_pdrop", "resid_pdrop") if self.backbone_spec.kind == "gpt2" else ("attention_dropout",)
        return max(float(getattr(cfg, k, 0.0) or 0.0) for k in keys) > 0

    def _check_input(self, inputs):
        x_enc = inputs["x_enc"]
        if not x_enc.is_cuda:
            raise MtsError("medtsllm_b200 runs on a CUDA device only (no CPU fallback)")
        if self._backbone is None:
            raise MtsError("model has not been moved to a CUDA device yet (call .to('cuda'))")
        if self.device is None:
            self.device = x_enc.device
        if x_enc.ndim == 2:
            x_enc = x_enc.unsqueeze(-1)
        if x_enc.dtype != torch.float32:
            raise MtsError(f"x_enc must be fp32 (setup.dtype 'mixed'/'float32'), got {x_enc.dtype}")
        B, T, C = x_enc.shape
        assert C == self.n_features and T == self.seq_len
        return x_enc.contiguous()

    def _forward_impl(self, inputs, stash, ids=None):
        """The hot path (models/medtsllm.py:321-382), everything before the eval-only activation.
        `stash` (dict) collects what train.backward needs; None for inference.  `ids`: the host prompt-id
        table when the caller has already built it."""
        x_enc = self._check_input(inputs)
        B, T, C = x_enc.shape
        bb = self._backbone
        dev = x_enc.device
        D, N, E, H = self.d_llm, self.n_patches, self.d_ff, self.n_attention_heads
        HE = H * E
        rl = self.reprogramming_layer
        mode = self.covariate_mode
        per_feature = mode not in ("concat", "univariate")      # features encoded separately: [B*C, N0, 32]
        N0 = N // C if mode == "interleave" else N               # patches per feature
        Bp = B * C if mode in ("independent", "merge-end") else B   # sequences through the backbone

        # K5: prompt ids (host) -> backbone input rows [0, Lp); repeated per feature for independent / merge-end
        if ids is None:
            ids = self.prompt_token_ids(inputs)
        Lp = ids.shape[1]
        L = Lp + N
        ids_dev = self._ids_device(ids, dev) if Lp > 0 else None
        # shared-prefix row layout: Lc leading prompt positions once, then Ls = L - Lc own rows per sequence
        # (not with live backbone dropouts: their masks differ per sample on the prompt rows too)
        bb_drop = self._backbone_dropout()
        self._last_backbone_dropout = bb_drop
        Lc = 0 if bb_drop is not None else self._shared_prefix_len(ids, Bp, L)
        Ls = L - Lc
        X = torch.empty(Lc + Bp * Ls, D, device=dev, dtype=torch.float32)
        ops.prompt_gather(ids_dev, bb.embed, bb.wpe, X, rep=Bp // B, Lp=Lp, L=L, Lc=Lc, B=B)

        # K1+K2: RevIN + patches + token conv (concat layout [B, N, C*32] or per feature [B*C, N0, 32])
        concat = mode == "concat"
        enc, _, mean, std = ops.revin_patch_embed(
            x_enc, self.patch_embedding.value_embedding.tokenConv.weight.detach(), self.patch_len, self.stride,
            concat=concat)
        assert enc.shape[1] == N0
        # train-mode dropout on the patch embeddings and the reprogramming attention (models/layers/embed.py:197,
        # models/medtsllm.py:587); seeds come from torch's CPU generator so torch.manual_seed reproduces a run
        p_drop = self._dropout_requested if self.training else 0.0
        seeds = [int(v) for v in torch.randint(0, 2 ** 62, (3,))] if p_drop > 0 else None
        self._last_dropout_seeds = seeds[:2] if seeds is not None else None
        if p_drop > 0:
            ops.dropout(enc, p_drop, seeds[0], out=enc)

        # time-series example parts of the prompt (`encode_part`'s tensor branch, models/medtsllm.py:313-319): each goes
        # through encode_ts like the window itself — own RevIN statistics, patches, token conv — and its patches join
        # the reprogramming rows of the batch (the layer is row-wise); the out-projection writes them to their prompt rows
        rows_main = enc.shape[0] * N0
        segs = getattr(ids, "example_segments", None) or []
        ex_meta = []
        enc_all = enc.view(rows_main, -1)
        if segs:
            conv_w = self.patch_embedding.value_embedding.tokenConv.weight.detach()
            ex_encs, r0 = [], rows_main
            for (b, pos, ts) in segs:
                ts = ts.to(dev, torch.float32).contiguous()
                e, _, m_e, s_e = ops.revin_patch_embed(ts, conv_w, self.patch_len, self.stride, concat=concat)
                ex_encs.append(e.view(e.shape[1], -1))
                ex_meta.append(dict(b=b, pos=pos, n=e.shape[1], r0=r0, ts=ts, mean=m_e, std=s_e))
                r0 += e.shape[1]
            ex_cat = torch.cat(ex_encs)
            if p_drop > 0:
                ops.dropout(ex_cat, p_drop, seeds[2], out=ex_cat)
            enc_all = torch.cat([enc_all, ex_cat])

        # K3/K4: reprogramming cross-attention on tcgen05 GEMMs
        source, K, Vt = self._source_kv()
        S = self.num_tokens
        rows = enc_all.shape[0]
        wq = self._bf16_weight("wq", rl.query_projection.weight)
        Q = torch.empty(rows, HE, device=dev, dtype=torch.bfloat16)
        ops.gemm(enc_all, wq, Q, m=rows, n=HE, k=self.d_model, ldb=wq.shape[1],
                 bias=rl.query_projection.bias.detach(), bias_axis=BIAS_N)
        scores = torch.empty(H, rows, S, device=dev, dtype=torch.float32)
        ops.gemm(Q, K, scores, m=rows, n=S, k=E, batch=H, lda=HE, ldb=HE, a_bs=E, b_bs=E, d_bs=rows * S)
        scale = 1.0 / math.sqrt(E)
        P = ops.softmax_rows(scores, scale)
        Pd = ops.dropout(P, p_drop, seeds[1]) if p_drop > 0 else P        # attention actually applied
        O = torch.empty(rows, HE, device=dev, dtype=torch.bfloat16)
        ops.gemm(Pd, Vt, O, m=rows, n=E, k=S, batch=H, a_bs=rows * S, ldb=S, b_bs=E * S, ldd=HE, d_bs=E)
        wo = self._bf16_weight("wo", rl.out_projection.weight)
        bo = rl.out_projection.bias.detach()
        Y = None
        if mode in ("concat", "univariate", "independent", "merge-end"):
            # one sequence per (sample[, feature]): rows of batch i land at X[i, Lp:, :]
            ops.gemm(O, wo, X, m=N, n=D, k=HE, batch=Bp, a_bs=N * HE, b_bs=0, d_bs=Ls * D, ldd=D, d_off=Lp * D,
                     ldb=wo.shape[1], bias=bo, bias_axis=BIAS_N, epilogue=EPI_RESID_ADD)
            for ex in ex_meta:      # example patches -> their rows inside the prompt (zero + wpe from the gather)
                ops.gemm(O, wo, X, m=ex["n"], n=D, k=HE, a_off=ex["r0"] * HE, ldd=D,
                         d_off=(Lc + ex["b"] * Ls + ex["pos"] - Lc) * D, ldb=wo.shape[1], bias=bo, bias_axis=BIAS_N,
                         epilogue=EPI_RESID_ADD)
        elif mode == "interleave":
            # token order n-major, c-minor (models/medtsllm.py:292-295): feature c writes rows Lp + n*C + c
            for c in range(C):
                ops.gemm(O, wo, X, m=N0, n=D, k=HE, batch=B, a_off=c * N0 * HE, a_bs=C * N0 * HE, b_bs=0,
                         d_bs=Ls * D, ldd=C * D, d_off=(Lp + c) * D, ldb=wo.shape[1], bias=bo, bias_axis=BIAS_N,
                         epilogue=EPI_RESID_ADD)
        else:
            # add / weighted-average (models/medtsllm.py:284-291): merge the C reprogrammed streams into one
            Y = torch.empty(rows, D, device=dev, dtype=torch.float32)
            ops.gemm(O, wo, Y, m=rows, n=D, k=HE, ldb=wo.shape[1], bias=bo, bias_axis=BIAS_N)
            fw = self.feature_weighting if mode == "weighted-average" else None
            ops.group_reduce(Y, B, C, N0 * D, w=fw.weight.detach().view(-1) if fw is not None else None,
                             bias=fw.bias.detach() if fw is not None else None, out=X, out_bs=Ls * D, out_off=Lp * D,
                             accumulate=True)       # X's patch rows hold 0 (+ wpe for GPT-2)

        cap = self._capture
        if cap is not None:
            cap.update(revin_mean=mean.clone(), revin_stdev=std.clone(), patch_embedding=enc.clone(),
                       source_embeddings=source.clone(), llm_input=self._expand_rows(X, Bp, L, Lc))
        # backbone
        layer_stash = [] if stash is not None else None
        hid, x_final = bb.forward(X, Bp, L, stash=layer_stash, lora=self.llm if self.lora_enabled else None,
                                  Lc=Lc, dropout=bb_drop)                         # bf16 rows like X, final norm applied
        if cap is not None:
            cap["llm"] = self._expand_rows(hid, Bp, L, Lc)
            cap["shared_prefix"] = Lc

        # K11: last N tokens -> Linear(D -> d_ff), stored transposed as [Bp, d_ff, N] (flatten index f*N+n)
        wds, bds = self._downsample_operands()
        flat = torch.empty(Bp, E * N, device=dev, dtype=torch.bfloat16)
        ops.gemm(hid, wds, flat, m=N, n=E, k=D, batch=Bp, a_off=Lp * D, a_bs=Ls * D, b_bs=0, d_bs=E * N,
                 d_transposed=True, ldd=N, ldb=wds.shape[1], bias=bds, bias_axis=BIAS_N if bds is not None else 0)
        # K12: flatten head
        wh = self._bf16_weight("wh", self.output_projection.linear.weight)
        head = torch.empty(Bp, self.n_outputs, device=dev, dtype=torch.float32)
        if (E * N) % 8:
            raise MtsError("d_ff * n_patches must be a multiple of 8")
        ops.gemm(flat, wh, head, m=Bp, n=self.n_outputs, k=E * N, ldb=wh.shape[1],
                 bias=self.output_projection.linear.bias.detach(), bias_axis=BIAS_N)
        if cap is not None:
            cap["output_projection"] = head.clone()
        nops = self.n_outputs_per_step
        if mode == "independent":          # mean over the feature axis (models/medtsllm.py:369-371)
            out = ops.group_reduce(head, B, C, self.n_outputs)
        elif mode == "merge-end":          # Linear over (feature, output) pairs (models/medtsllm.py:372-375)
            out = ops.merge_end(head, self.feature_weighting.weight.detach(), self.feature_weighting.bias.detach(),
                                B, C, self.pred_len, nops)
        else:
            out = head
        out = out.view(B, self.pred_len, nops)
        denorm = self.task in ("forecasting", "reconstruction", "anomaly_detection", "pretraining")
        if denorm:
            ops.revin_denorm(out, mean, std)
        else:
            out = out.squeeze(-1)
        if stash is not None:
            stash.update(x_enc=x_enc, mean=mean, std=std, enc=enc, source=source, K=K, Vt=Vt, Q=Q, P=P, O=O,
                         hid=hid, x_final=x_final, flat=flat, layers=layer_stash, Lp=Lp, L=L, Lc=Lc, Bp=Bp, B=B, N0=N0,
                         scale=scale, denorm=denorm, concat=concat, Y=Y, head=head, Pd=Pd, p_drop=p_drop, seeds=seeds,
                         bb_drop=bb_drop, enc_all=enc_all, ex_meta=ex_meta, rows_main=rows_main)
        return out


def _calcute_lags(x, n_lags=5):
    """models/medtsllm.py:530-538 (prompt text only)."""
    x = x.permute(0, 2, 1).contiguous() if x.ndim == 3 else x.unsqueeze(1)
    q_fft = torch.fft.rfft(x, dim=-1)
    res = q_fft * torch.conj(q_fft)
    corr = torch.fft.irfft(res, dim=-1)
    mean_value = torch.mean(corr, dim=1)
    _, lags = torch.topk(mean_value, n_lags, dim=-1)
    return lags
