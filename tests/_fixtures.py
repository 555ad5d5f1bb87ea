"""Helpers shared by the parity tests: load a golden fixture (tests/golden/*.pt, produced by
oracle/make_golden.py from the unmodified reference), rebuild its HF backbone directory, and run the
oracle restatement on it."""
from __future__ import annotations

import copy
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
for p in (REPO, REPO / "med-ts-llm_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

CASES = ["llama_seg_concat", "gpt2_anomaly_concat", "llama_semseg_univariate", "llama_forecast_clip_stats",
         "llama_forecast_truncate", "gpt2_reconstruction_average", "llama_forecast_independent",
         "gpt2_forecast_merge_end", "llama_anomaly_add", "gpt2_anomaly_weighted_average", "llama_forecast_interleave",
         "llama_seg_examples"]


def same_items(a, b) -> bool:
    """Equality of prompt parts / prompt items that may hold time-series tensors (prompting.examples)."""
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(same_items(x, y) for x, y in zip(a, b))
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and torch.equal(a.cpu(), b.cpu())
    return a == b


def prompt_items(model, inputs):
    """The kernel model's host prompt table as the oracle wants it: per sample a flat list of token ids (left padding
    included) with every run of example-part positions (negative ids) replaced by the example tensor itself."""
    table = model.prompt_token_ids(inputs)
    segs = {(b, pos): ts for (b, pos, ts) in (getattr(table, "example_segments", None) or [])}
    out = []
    for b, row in enumerate(table.tolist()):
        items, pos = [], 0
        while pos < len(row):
            ts = segs.get((b, pos))
            if ts is not None:
                items.append(ts.detach().cpu())
                n = 0
                while pos + n < len(row) and row[pos + n] < 0 and (n == 0 or (b, pos + n) not in segs):
                    n += 1
                pos += n
            else:
                assert row[pos] >= 0
                items.append(row[pos])
                pos += 1
        out.append(items)
    return out


def load_case(name: str) -> dict:
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


class Dataset:
    def __init__(self, d):
        self.n_features = d["n_features"]
        self.n_classes = d["n_classes"]
        self.description = d["description"]
        self.task_description = None


def hf_model_from_fixture(fix):
    import transformers
    cfg_dict = dict(fix["hf_config"])
    cfg_dict.pop("model_type", None)
    cfg_dict.pop("transformers_version", None)
    if fix["kind"] == "llama":
        config = transformers.LlamaConfig(**cfg_dict)
        model = transformers.LlamaModel(config)
    else:
        config = transformers.GPT2Config(**cfg_dict)
        model = transformers.GPT2Model(config)
    missing = model.load_state_dict({k: v.float() for k, v in fix["backbone_state"].items()}, strict=False)
    assert not missing.unexpected_keys, missing
    return model.eval()


def materialize_llm_dir(fix, path: Path) -> Path:
    """Writes config + weights + tokenizer so that AutoModel/AutoTokenizer load it like a checkpoint."""
    from tokenizers import Tokenizer
    from transformers import PreTrainedTokenizerFast
    path.mkdir(parents=True, exist_ok=True)
    hf_model_from_fixture(fix).save_pretrained(str(path))
    tok = Tokenizer.from_str(fix["tokenizer_json"])
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>",
                                   bos_token="<s>" if fix["tokenizer_bos"] else None, eos_token="</s>")
    fast.save_pretrained(str(path))
    return path


def config_for(fix, llm_dir):
    cfg = copy.deepcopy(fix["config"])
    cfg["models"]["medtsllm"]["llm"]["llm"] = str(llm_dir)
    return cfg


class Cfg(dict):
    """Minimal attribute-dict with the surface of the reference's dict_to_object (utils.py:19-39)."""

    def __init__(self, d):
        super().__init__({k: Cfg(v) if isinstance(v, dict) else v for k, v in d.items()})

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def oracle_spec(fix) -> dict:
    cfg = fix["config"]
    mc = cfg["models"]["medtsllm"]
    hc = fix["hf_config"]
    task = cfg["task"]
    C = fix["dataset"]["n_features"]
    ncls = fix["dataset"]["n_classes"] if task == "semantic_segmentation" else 0
    if task in ("forecasting", "reconstruction", "anomaly_detection", "pretraining"):
        nops = C
    elif task == "semantic_segmentation":
        nops = ncls if ncls > 2 else 1
    else:
        nops = 1
    llama = fix["kind"] == "llama"
    return dict(
        task=task, pred_len=cfg["pred_len"], patch_len=mc["patching"]["patch_len"], stride=mc["patching"]["stride"],
        d_model=mc["d_model"], d_ff=mc["d_ff"], n_heads=mc["n_heads"], covariate_mode=mc["covariate_mode"],
        downsample=mc["embedding_downsample_mode"], n_outputs_per_step=nops, backbone=fix["kind"],
        n_layers=hc["num_hidden_layers"] if llama else hc["n_layer"],
        llm_heads=hc["num_attention_heads"] if llama else hc["n_head"],
        eps=hc["rms_norm_eps"] if llama else hc["layer_norm_epsilon"],
        rope_theta=(hc.get("rope_parameters") or {}).get("rope_theta", 10000.0) if llama else None,
        pad_id=fix["pad_id"], seg_mode=cfg["tasks"]["segmentation"]["mode"], n_classes=ncls)


def run_oracle(fix, training=False, prompt_ids=None):
    from oracle import medtsllm_oracle as O
    sd = {k: v.float() for k, v in fix["backbone_state"].items()}
    return O.medtsllm_forward(fix["inputs"]["x_enc"], prompt_ids or fix["prompt_ids"], fix["adapters"], sd,
                              oracle_spec(fix), training=training, return_stages=True)


# ------------------------------------------------------------------------------------------------ GPT4TS
GPT4TS_CASES = ["gpt4ts_forecast_etth1", "gpt4ts_anomaly", "gpt4ts_semseg", "gpt4ts_segmentation"]


def load_gpt4ts_backbone():
    """Backbone shared by the GPT4TS fixtures (tests/golden/gpt4ts_backbone.pt, made by
    oracle/make_golden_gpt4ts.py): HF config + the state of the blocks GPT4TS keeps, bf16-representable."""
    return torch.load(GOLDEN / "gpt4ts_backbone.pt", weights_only=False)


def gpt4ts_spec(fix, bb) -> dict:
    cfg = fix["config"]
    return dict(task=cfg["task"], pred_len=cfg["pred_len"], d_ff=cfg["models"]["gpt4ts"]["d_ff"],
                gpt_layers=cfg["models"]["gpt4ts"]["gpt_layers"], n_heads=bb["hf_config"]["n_head"],
                eps=bb["hf_config"]["layer_norm_epsilon"], n_classes=fix["dataset"]["n_classes"],
                seg_mode=cfg["tasks"]["segmentation"]["mode"])


def gpt4ts_hf_model(bb, n_layers):
    """HF GPT2Model holding the fixture backbone (wte is unused by GPT4TS and not stored: left at its random init)."""
    import transformers
    d = dict(bb["hf_config"])
    d.pop("model_type", None); d.pop("transformers_version", None)
    d["n_layer"] = n_layers
    model = transformers.GPT2Model(transformers.GPT2Config(**d))
    res = model.load_state_dict({k: v.float() for k, v in bb["state"].items()}, strict=False)
    assert not res.unexpected_keys and all(k.startswith("wte") for k in res.missing_keys), res
    return model.eval()
