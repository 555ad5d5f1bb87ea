"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference (flixpar/med-ts-llm at /root/reference) on CPU.

Only usable in the build container (the GPU box has no /root/reference).  It exists to
  (1) pin oracle/medtsllm_oracle.py (the travelling CPU restatement) against the reference itself,
  (2) generate the golden fixtures under tests/golden/ (see oracle/make_golden.py).

The reference is imported as-is; the shims below only stand in for third-party packages that are
absent from this image and never touch reference source:
  - `peft`            (models/medtsllm.py:12-13)  stubbed during import, then removed
  - `matplotlib`, `reformer_pytorch`  (eager baseline imports, models/__init__.py:5-7)
  - `AutoModel` inside models.medtsllm: drops `device_map="auto"` (needs `accelerate`, :183)
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types
from pathlib import Path

import torch

REFERENCE = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE / "models" / "medtsllm.py").exists()


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_ref = None


def import_reference():
    """Returns a namespace with the reference's MedTsLLM class and dict_to_object."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("/root/reference is not present (this only runs in the build container)")
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    # the reference's top-level packages have generic names; make sure ours do not shadow them
    for name in ("models", "utils", "datasets", "tasks", "loggers"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(str(REFERENCE)):
            del sys.modules[name]

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    added = []
    if "peft" not in sys.modules:
        _stub("peft", LoraConfig=_Dummy, TaskType=types.SimpleNamespace(FEATURE_EXTRACTION="FE"),
              get_peft_model=lambda m, c: m)
        added.append("peft")
    for name in ("matplotlib", "matplotlib.pyplot", "reformer_pytorch"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name, LSHSelfAttention=_Dummy)
                added.append(name)
    if "matplotlib" in added:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    medtsllm = importlib.import_module("models.medtsllm")
    utils = importlib.import_module("utils")
    for name in added:  # HF's is_peft_available() chokes on spec-less stubs
        sys.modules.pop(name, None)

    real_auto_model = medtsllm.AutoModel

    class _AutoModelNoDeviceMap:
        @staticmethod
        def from_pretrained(*args, **kwargs):
            kwargs.pop("device_map", None)
            return real_auto_model.from_pretrained(*args, **kwargs)

    medtsllm.AutoModel = _AutoModelNoDeviceMap
    _ref = types.SimpleNamespace(MedTsLLM=medtsllm.MedTsLLM, dict_to_object=utils.dict_to_object,
                                 module=medtsllm)
    return _ref


# --------------------------------------------------------------------------------------------------
# Synthetic backbone + tokenizer (no checkpoints exist offline)
# --------------------------------------------------------------------------------------------------
def build_tokenizer(texts, vocab_size: int, out_dir: Path, bos: bool):
    """Deterministic WordLevel tokenizer over the words of `texts` (+ digits/punctuation)."""
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast

    specials = ["<unk>", "<s>", "</s>"]
    words = []
    for t in texts:
        for w in pre_tokenizers.Whitespace().pre_tokenize_str(t):
            if w[0] not in words:
                words.append(w[0])
    for ch in "0123456789.-,[]=():":
        if ch not in words:
            words.append(ch)
    vocab = {w: i for i, w in enumerate(specials + words)}
    if len(vocab) > vocab_size:
        raise ValueError(f"vocab_size {vocab_size} too small for {len(vocab)} words")
    i = 0
    while len(vocab) < vocab_size:
        vocab[f"<extra_{i}>"] = len(vocab)
        i += 1
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    if bos:  # Llama-style: every encoded string gets a BOS (add_special_tokens default True)
        tok.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", bos_token="<s>" if bos else None,
                                   eos_token="</s>")
    fast.save_pretrained(str(out_dir))
    return fast


def build_llm_dir(kind: str, out_dir: Path, texts, *, seed: int = 0, **cfg):
    """Creates a local HF model directory (random init, seeded) + tokenizer.  kind: llama | gpt2."""
    import transformers
    out_dir.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(seed)
    if kind == "llama":
        config = transformers.LlamaConfig(
            vocab_size=cfg.get("vocab_size", 512), hidden_size=cfg.get("hidden_size", 128),
            intermediate_size=cfg.get("intermediate_size", 256), num_hidden_layers=cfg.get("layers", 2),
            num_attention_heads=cfg.get("heads", 2), num_key_value_heads=cfg.get("heads", 2),
            rms_norm_eps=1e-5, max_position_embeddings=cfg.get("max_pos", 512),
            bos_token_id=1, eos_token_id=2, tie_word_embeddings=False)
        model = transformers.LlamaModel(config)
    elif kind == "gpt2":
        config = transformers.GPT2Config(
            vocab_size=cfg.get("vocab_size", 512), n_embd=cfg.get("hidden_size", 128),
            n_layer=cfg.get("layers", 2), n_head=cfg.get("heads", 2), n_positions=cfg.get("max_pos", 512),
            bos_token_id=2, eos_token_id=2,
            # the real checkpoints carry 0.1 dropouts that stay LIVE in the reference's train mode
            # (model.train() flips the frozen backbone too); fixtures need a deterministic train pass
            attn_pdrop=cfg.get("pdrop", 0.0), embd_pdrop=cfg.get("pdrop", 0.0), resid_pdrop=cfg.get("pdrop", 0.0))
        model = transformers.GPT2Model(config)
    else:
        raise ValueError(kind)
    model.save_pretrained(str(out_dir))
    build_tokenizer(texts, config.vocab_size, out_dir, bos=(kind == "llama"))
    return out_dir


# --------------------------------------------------------------------------------------------------
# Config + dataset stand-ins with exactly the attributes MedTsLLM.__init__ reads
# --------------------------------------------------------------------------------------------------
class SyntheticDataset:
    """Attributes read by the reference ctor: models/medtsllm.py:41-42,56,58."""

    def __init__(self, n_features, n_classes=0, description="", task_description=None):
        self.n_features = n_features
        self.n_classes = n_classes
        self.description = description
        self.task_description = task_description


def make_config(*, task, history_len, pred_len, llm_path, d_model=32, d_ff=64, n_heads=8,
                num_tokens=1024, covariate_mode="concat", downsample="linear", patch_len=16, stride=8,
                dropout=0.0, dtype="float32", llm_layers=-1, prompting=None, seg_mode="boundary-prediction"):
    prompting = prompting or {"dataset": True, "task": True, "clip": False, "input_stats": False,
                              "examples": False, "input_stats_dim": 0, "input_stats_select": "all"}
    return {
        "task": task, "model": "medtsllm", "history_len": history_len, "pred_len": pred_len,
        "training": {"dropout": dropout, "batch_size": 4, "learning_rate": 1e-4},
        "setup": {"dtype": dtype, "seed": 0},
        "tasks": {"segmentation": {"mode": seg_mode}},
        "models": {"medtsllm": {
            "d_model": d_model, "d_ff": d_ff, "n_heads": n_heads, "num_tokens": num_tokens,
            "covariate_mode": covariate_mode, "embedding_downsample_mode": downsample,
            "patching": {"patch_len": patch_len, "stride": stride},
            "prompting": prompting,
            "llm": {"enabled": True, "llm": str(llm_path), "llm_layers": llm_layers,
                    "load_in_4bit": False, "load_in_8bit": False},
        }},
    }


@contextlib.contextmanager
def _quiet():
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def build_reference_model(config: dict, dataset: SyntheticDataset, seed: int = 0):
    ref = import_reference()
    torch.manual_seed(seed)
    with _quiet():
        model = ref.MedTsLLM(ref.dict_to_object(config), dataset)
    return model


def run_reference_with_stages(model, inputs: dict, train_mode: bool = False):
    """Forward through the reference, capturing the per-stage tensors the parity tests compare."""
    stages = {}
    hooks = []

    def grab(name):
        def fn(_m, _inp, out):
            o = out[0] if isinstance(out, tuple) else out
            if hasattr(o, "last_hidden_state"):
                stages[name + ".hidden_states"] = [h.detach().clone() for h in o.hidden_states]
                o = o.last_hidden_state
            stages[name] = o.detach().clone()
        return fn

    for name in ("patch_embedding", "mapping_layer", "reprogramming_layer", "llm", "output_projection"):
        hooks.append(getattr(model, name).register_forward_hook(grab(name)))
    if hasattr(model, "embedding_downsample_layer"):
        hooks.append(model.embedding_downsample_layer.register_forward_hook(grab("embedding_downsample_layer")))
    # the LLM input (prompt + reprogrammed patches)
    hooks.append(model.llm.register_forward_pre_hook(
        lambda _m, args, kwargs: stages.__setitem__("llm_input", kwargs["inputs_embeds"].detach().clone()),
        with_kwargs=True))
    model.train(train_mode)
    try:
        with torch.set_grad_enabled(train_mode):
            out = model(inputs)
    finally:
        for h in hooks:
            h.remove()
    stages["revin_mean"] = model.normalize_layers.mean.detach().clone()
    stages["revin_stdev"] = model.normalize_layers.stdev.detach().clone()
    stages["output"] = out.detach().clone()
    return out, stages


# --------------------------------------------------------------------------------------------------
# The reference's Trainer (tasks/*, datasets/*, loggers/*), imported unmodified
# --------------------------------------------------------------------------------------------------
_trainer_ns = None


def import_trainer():
    """Imports the reference's `tasks`, `datasets` and `loggers` packages as they are.  In-memory stubs stand in only
    for third-party packages that are absent from this image and that the Trainer imports at module level without
    using them on the paths exercised here: `pytorch_optimizer` (tasks/base.py:12, Ranger21 only), `bayes_opt`
    (tasks/anomaly_detection.py:14), `plotly.graph_objects` (:16), `wfdb` (datasets/ludb.py)."""
    global _trainer_ns
    if _trainer_ns is not None:
        return _trainer_ns
    ref = import_reference()

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    for name, attrs in (("pytorch_optimizer", {"Ranger21": _Dummy}), ("bayes_opt", {"BayesianOptimization": _Dummy}),
                        ("plotly", {}), ("plotly.graph_objects", {"Figure": _Dummy}), ("wfdb", {})):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name, **attrs)
    if "plotly" in sys.modules and not hasattr(sys.modules["plotly"], "graph_objects"):
        sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    # the reference's top-level `datasets` package must win over HuggingFace `datasets` if that was imported earlier
    for name in ("datasets", "tasks", "loggers"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(str(REFERENCE)):
            del sys.modules[name]
    datasets = importlib.import_module("datasets")
    tasks = importlib.import_module("tasks")
    loggers = importlib.import_module("loggers")
    tasks_base = importlib.import_module("tasks.base")
    datasets_base = importlib.import_module("datasets.base")
    _trainer_ns = types.SimpleNamespace(ref=ref, datasets=datasets, tasks=tasks, loggers=loggers,
                                        tasks_base=tasks_base, datasets_base=datasets_base,
                                        dict_to_object=ref.dict_to_object)
    return _trainer_ns
