"""CPU-only tests of the host side of the boundary: the C ABI surface, the drop-in class surface
(constructor, state_dict keys, prompt building/tokenisation) and loud failure without a GPU."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from _fixtures import CASES, Cfg, Dataset, config_for, load_case, materialize_llm_dir, same_items

REPO = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from medtsllm_b200 import _lib
    header = (REPO / "include" / "mts_b200.h").read_text()
    declared = set(re.findall(r"^\s*(?:int|const char\*|int64_t)\s+(mts_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared, "no prototypes parsed"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/mts_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.version() == 2
    assert ctypes.sizeof(_lib.GemmArgs) == 248      # layout of mts_gemm_args (8-byte aligned)


def test_process_wide_options_are_validated():
    """mts_set_option (include/mts_b200.h): known names with legal values are accepted without a device, anything else is
    an error with a message — a mistyped switch must not silently run the default schedule."""
    from medtsllm_b200 import MtsError, _lib
    for name, value in (("gemm_ksplit", 0), ("gemm_ksplit", 2), ("gemm_ksplit", 4), ("gemm_ksplit", -1), ("streamk", 3),
                        ("streamk", 0), ("epi_direct", 1), ("gemm_2cta", 1), ("pdl", 1), ("attn_tc", 1)):
        _lib.set_option(name, value)
    with pytest.raises(MtsError, match="gemm_ksplit"):
        _lib.set_option("gemm_ksplit", 3)
    with pytest.raises(MtsError, match="unknown option"):
        _lib.set_option("gemm_split_k", 2)


def test_no_cpu_fallback():
    from medtsllm_b200 import MtsError, ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(MtsError):
        ops.gemm(a, a, a, m=8, n=8, k=8)
    with pytest.raises(MtsError):
        ops.rmsnorm(torch.zeros(2, 8), torch.ones(8), 1e-5)
    if not torch.cuda.is_available():
        # with valid-looking arguments but no device the library itself must refuse
        from medtsllm_b200 import _lib
        with pytest.raises(MtsError):
            _lib.call("mts_cast_f32_bf16", 16, 32, 8, None)


@pytest.mark.parametrize("name", CASES)
def test_dropin_surface_and_prompts(name, tmp_path):
    from medtsllm_b200 import MtsError
    from medtsllm_b200.model import MedTsLLM
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    model = MedTsLLM(Cfg(config_for(fix, llm_dir)), Dataset(fix["dataset"]))
    # checkpoint contract: exactly the reference's adapter keys, same shapes
    sd = model.state_dict()
    assert list(sd.keys()) == list(fix["adapters"].keys())
    for k, v in fix["adapters"].items():
        assert sd[k].shape == v.shape, k
    res = model.load_state_dict(fix["adapters"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert trainable == set(fix["adapters"].keys())
    assert "forecasting" in model.supported_tasks and model.lora_enabled is False
    # prompt text + tokenisation identical to the reference's (models/medtsllm.py:386-439, :299-302)
    inputs = dict(fix["inputs"])
    assert same_items(model.build_prompt(inputs), fix["prompts"])
    table = model.prompt_token_ids(inputs)
    if "examples" in fix["inputs"]:
        # prompting.examples: a time-series part occupies n_patches(T_ex) positions marked by negative ids
        from _fixtures import prompt_items
        items = prompt_items(model, inputs)
        Lp = table.shape[1]
        assert Lp == fix["stages"]["llm_input"].shape[1] - model.n_patches        # the reference's own prompt length
        for b, want in enumerate(fix["prompt_ids"]):
            got = items[b]
            n_pad = len(got) - len(want)
            assert all(t == fix["pad_id"] for t in got[:n_pad]) and same_items(got[n_pad:], want)
            assert sorted(set(t for t in table[b].tolist() if t < 0)) == [-(b + 1)]
    else:
        Lp = max(len(p) for p in fix["prompt_ids"])
        assert table.shape == (len(fix["prompt_ids"]), Lp) and table.dtype == torch.int32
        for b, ids in enumerate(fix["prompt_ids"]):
            assert table[b, Lp - len(ids):].tolist() == ids
            assert all(t == fix["pad_id"] for t in table[b, : Lp - len(ids)].tolist())
    # load_pretrained drops the head (models/medtsllm.py:515-527)
    loaded = model.load_pretrained(dict(fix["adapters"]))
    assert "output_projection.linear.weight" not in loaded and "mapping_layer.weight" in loaded
    # no silent CPU path
    with pytest.raises(MtsError):
        model.eval()(inputs)


def test_unsupported_options_fail_loudly(tmp_path):
    from medtsllm_b200.model import MedTsLLM
    fix = load_case("llama_seg_concat")
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    cfg = config_for(fix, llm_dir)
    cfg["models"]["medtsllm"]["covariate_mode"] = "stacked"
    with pytest.raises(ValueError):
        MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    cfg["models"]["medtsllm"]["covariate_mode"] = "independent"
    cfg["models"]["medtsllm"]["prompting"]["examples"] = True
    ex = {**fix["inputs"], "examples": [("Example segment:", torch.zeros(1, 40, 3))] * fix["inputs"]["x_enc"].shape[0]}
    with pytest.raises(NotImplementedError):     # the reference's own torch.cat of prompt parts breaks for this mode
        MedTsLLM(Cfg(cfg), Dataset(fix["dataset"])).prompt_token_ids(ex)
    cfg = config_for(fix, llm_dir)
    cfg["models"]["medtsllm"]["lora"] = {"enabled": True, "layers": "auto", "rank": 8, "alpha": 16, "dropout": 0.1}
    with pytest.raises(NotImplementedError):
        MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    # LoRA itself is supported: adapters live under model.llm and stay out of the checkpoint like the reference
    cfg["models"]["medtsllm"]["lora"]["dropout"] = 0.0
    m = MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    assert m.lora_enabled and not any(k.startswith("llm.") for k in m.state_dict())
    n_lora = sum(p.numel() for n, p in m.named_parameters() if n.startswith("llm."))
    assert n_lora == 2 * 2 * 2 * 8 * 128          # layers x (q, v) x (A, B) x r x D


def test_plugin_registers_into_reference_model_lookup(tmp_path):
    """Drop-in seam (tasks/base.py:81-85): after register(), the reference's own lookup builds OUR class from
    the reference's own config object.  Needs the reference tree, i.e. runs in the build container only."""
    from oracle import ref_harness as H
    if not H.reference_available():
        pytest.skip("/root/reference not present on this machine")
    ref = H.import_reference()              # installs the third-party shims and imports `models`
    import models
    original = dict(models.model_lookup)
    try:
        from medtsllm_b200 import plugin
        from medtsllm_b200.model import MedTsLLM
        lookup = plugin.register()
        assert lookup["medtsllm"] is MedTsLLM and lookup["timellm"] is MedTsLLM
        fix = load_case("gpt2_anomaly_concat")
        llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
        cfg = ref.dict_to_object(config_for(fix, llm_dir))          # the reference's config type
        model = models.model_lookup[cfg.model](cfg, Dataset(fix["dataset"]))   # == BaseTask.build_model
        assert isinstance(model, MedTsLLM) and cfg.task in model.supported_tasks
        assert list(model.state_dict().keys()) == list(fix["adapters"].keys())
    finally:
        models.model_lookup.clear(); models.model_lookup.update(original)


def test_gpt4ts_surface_matches_reference_names():
    """Constructor / parameter names of medtsllm_b200.GPT4TS against the state captured from the unmodified
    models/gpt4ts.py (tests/golden/gpt4ts_*.pt); CPU tensors are refused (no fallback)."""
    import pytest
    from _fixtures import Cfg, Dataset, GPT4TS_CASES, load_case
    from medtsllm_b200._lib import MtsError
    from medtsllm_b200.backbone import BackboneSpec, KernelBackbone
    from medtsllm_b200.gpt4ts import GPT4TS
    bb = KernelBackbone(BackboneSpec("gpt2", 768, 2, 12, 768, 64, 1e-5), "cpu")     # never run: surface only
    for name in GPT4TS_CASES:
        fix = load_case(name)
        model = GPT4TS(Cfg(fix["config"]), Dataset(fix["dataset"]), backbone=bb)
        own = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.endswith("position_embedding.pe")}
        assert own == {k: tuple(v.shape) for k, v in fix["params"].items()}, name
        with pytest.raises(MtsError):
            model({"x_enc": fix["inputs"]["x_enc"]})
    cfg = dict(load_case("gpt4ts_anomaly")["config"]); cfg["task"] = "reconstruction"
    with pytest.raises(ValueError):      # models/gpt4ts.py:103-104: listed in supported_tasks, rejected by forward
        GPT4TS(Cfg(cfg), Dataset(fix["dataset"]), backbone=bb)({"x_enc": fix["inputs"]["x_enc"]})


def test_lora_save_load_round_trip(tmp_path):
    """`model.llm.save_pretrained` (what loggers/base_logger.py:42-43 calls) and its inverse."""
    import torch
    from medtsllm_b200.backbone import BackboneSpec
    from medtsllm_b200.lora import LoraAdapters
    spec = BackboneSpec("llama", 64, 2, 2, 128, 100, 1e-5)
    a = LoraAdapters(spec, rank=4, alpha=8, rslora=True, init=True)
    with torch.no_grad():
        for p in a.params():
            p.copy_(torch.randn(p.shape))
    path = tmp_path / "run" / "latest-lora.safetensors"
    a.save_pretrained(path)
    b = LoraAdapters(spec, rank=4, alpha=8, rslora=True, init=True)
    v0 = [p._version for p in b.params()]
    keys = b.load_pretrained(path)
    assert len(keys) == len(a.params())
    for pa, pb, v in zip(a.params(), b.params(), v0):
        assert torch.equal(pa, pb) and pb._version > v


def test_shared_prefix_length_host_logic(tmp_path):
    """`MedTsLLM._shared_prefix_len`: number of leading prompt positions (left padding included) holding the same token
    in every sample — the rows carried once through the backbone.  Host logic only."""
    from _fixtures import Cfg, Dataset, config_for, load_case, materialize_llm_dir
    from medtsllm_b200.model import MedTsLLM
    fix = load_case("llama_seg_concat")
    model = MedTsLLM(Cfg(config_for(fix, materialize_llm_dir(fix, tmp_path / "llm"))), Dataset(fix["dataset"]))
    B, Lp, N = 6, 40, 13
    L = Lp + N
    same = torch.randint(3, 300, (1, Lp), dtype=torch.int32).repeat(B, 1)
    assert model._shared_prefix_len(same, B, L) == Lp                    # static dataset / task prompt: all of it
    t = same.clone(); t[3, 25] += 1
    assert model._shared_prefix_len(t, B, L) == 25                       # per-sample text from position 25 on
    t = same.clone(); t[1, 9] += 1
    assert model._shared_prefix_len(t, B, L) == 0                        # fewer than 16 shared positions: not worth a launch
    t = same.clone(); t[2, :4] = 2; t[2, 4:] = same[0, : Lp - 4]         # a shorter prompt, left-padded: nothing lines up
    assert model._shared_prefix_len(t, B, L) == 0
    assert model._shared_prefix_len(same, 1, L) == 0                     # a single sequence has nothing to share with
    assert model._shared_prefix_len(same[:, :0], B, N) == 0              # no prompt
    model.share_prompt_prefix = False
    assert model._shared_prefix_len(same, B, L) == 0
    model.share_prompt_prefix = True
    # the sequence-resident attention kernels must hold all L positions of one head in shared memory
    long = torch.randint(3, 300, (1, 600), dtype=torch.int32).repeat(B, 1)
    assert model._shared_prefix_len(long, B, 600 + N) == 0
