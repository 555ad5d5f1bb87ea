"""Whole-path parity on the GPU: medtsllm_b200.MedTsLLM (CUDA kernels through the C ABI) against
(a) the oracle restatement run on CPU on the same inputs and (b) the golden tensors captured from the
unmodified reference (tests/golden/*.pt).

Tolerance (stated, per SURVEY.md §7 "hard parts"): the kernels compute GEMM operands in bf16 with
fp32 accumulation and an fp32 residual stream — the reference's own GPU regime under bf16 autocast
(tasks/forecasting.py:22) — while oracle/goldens are fp32.  Metric: relative L2.
  front end (fp32 math, bf16 store) ............ < 4e-3
  intermediate stages (pre-activation) ......... < 1e-2   (measured 0.7e-3 .. 4.9e-3)
  final output ................................. < 3e-3   (measured 2.8e-4 .. 1.5e-3)
Yardstick (oracle/yardstick.py, run on the unmodified reference): the reference's OWN bf16-autocast forward
differs from its fp32 forward by 0.8e-3 .. 2.2e-3 on these cases, i.e. the kernel path sits inside the
reference's own mixed-precision noise.  north_star's 1e-3 is met by 9 of the 11 cases and missed by at most 1.5x.
The fixtures' backbone weights are bf16-representable, so weight rounding is not in this budget.
"""
import pytest
import torch

from _fixtures import CASES, Cfg, Dataset, config_for, load_case, materialize_llm_dir, prompt_items, run_oracle

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# per-precision tolerances: (intermediate stages, final output)
_TOL = {"bf16": (1e-2, 3e-3), "tf32": (2e-3, 1e-3), "fp32": (1e-4, 1e-4)}


@pytest.mark.parametrize("precision", ["bf16", "tf32", "fp32"])
@pytest.mark.parametrize("name", CASES)
def test_forward_parity(name, precision, tmp_path, cuda, monkeypatch):
    """precision "bf16": the default path.  "tf32" / "fp32": the evaluation parity modes (precise.py) — here the
    asserted output tolerance is north_star's 1e-3 (tf32: the reference's own evaluation regime) and 1e-4 (fp32)."""
    from medtsllm_b200.model import MedTsLLM
    monkeypatch.setenv("MTS_PRECISION", precision)
    tol_stage, tol_out = _TOL[precision]
    fix = load_case(name)
    if "examples" in fix["inputs"] and precision != "bf16":
        pytest.skip("prompting.examples is implemented on the bf16 path (the parity modes raise NotImplementedError)")
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    model = MedTsLLM(Cfg(config_for(fix, llm_dir)), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda, torch.float32).eval()
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    model._capture = {}
    with torch.no_grad():
        out = model(inputs)
    torch.cuda.synchronize()
    cap, model._capture = model._capture, None
    # the input-statistics prompt ranks FFT autocorrelation lags whose values tie exactly in theory
    # (corr[k] == corr[T-k]), so CPU and GPU can order them differently: embed the GPU-side tokens in
    # the oracle, and compare with the CPU-generated goldens only when the prompts agree
    ids = prompt_items(model, inputs)        # token ids; time-series example parts (prompting.examples) as tensors
    Lp_max = max(len(p) for p in fix["prompt_ids"])
    same_prompt = "examples" in fix["inputs"] or all(r[Lp_max - len(p):] == p for r, p in zip(ids, fix["prompt_ids"]))
    ref_out, st = run_oracle(fix, prompt_ids=ids)
    g = fix["stages"] if same_prompt else {**st, "output": ref_out, "output_train": run_oracle(fix, True, ids)[0]}

    assert out.shape == g["output"].shape and out.dtype == torch.float32
    B, T, C = fix["inputs"]["x_enc"].shape
    # front end
    torch.testing.assert_close(cap["revin_mean"].cpu().view(B, 1, C), g["revin_mean"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(cap["revin_stdev"].cpu().view(B, 1, C), g["revin_stdev"], rtol=1e-5, atol=1e-5)
    pe = st["patch_embedding"]
    if fix["config"]["models"]["medtsllm"]["covariate_mode"] == "concat":
        N = pe.shape[1]
        pe = pe.reshape(B, C, N, -1).permute(0, 2, 1, 3).reshape(B, N, -1)
    assert _rel_l2(cap["patch_embedding"], pe) < (4e-3 if precision == "bf16" else 1e-5)
    assert model.precision == precision
    # stages: vs oracle (run here) and vs the reference goldens
    if fix["kind"] == "gpt2":
        # the kernel path folds GPT-2's position embedding into the gather (HF adds wpe inside the model,
        # HF:models/gpt2/modeling_gpt2.py:584-585); the reference-side capture is pre-wpe
        L = cap["llm_input"].shape[1]
        cap["llm_input"] = cap["llm_input"] - model._backbone.wpe[:L][None]
    for key in ("source_embeddings", "llm_input", "llm", "output_projection"):
        e_o = _rel_l2(cap[key].float().view(st[key].shape), st[key])
        e_g = _rel_l2(cap[key].float().view(g[key].shape), g[key])
        assert e_o < tol_stage and e_g < tol_stage, (name, key, e_o, e_g)
    e_out = _rel_l2(out, g["output"])
    assert e_out < tol_out, (name, e_out)
    assert _rel_l2(out, ref_out) < tol_out
    # determinism: same inputs -> bit-identical output (no atomics on the forward path)
    with torch.no_grad():
        out2 = model(inputs)
    assert torch.equal(out, out2)
    # train()-mode forward under no_grad skips the eval-only activation (models/medtsllm.py:251-259)
    model.train()
    with torch.no_grad():
        out_t = model(inputs)
    assert _rel_l2(out_t, g["output_train"]) < 1e-2       # pre-activation (logits) in train mode
    print(f"\n[parity] {name} [{precision}]: rel-L2 output {e_out:.2e}  " +
          "  ".join(f"{k} {_rel_l2(cap[k].float().view(g[k].shape), g[k]):.1e}" for k in
                    ("source_embeddings", "llm_input", "llm", "output_projection")))


def test_random_init_backbone_matches_oracle_llama_hd128(cuda):
    """Device-generated random-init stack (what bench.py uses): pull its weights back and run the
    oracle's Llama restatement on them."""
    from medtsllm_b200.backbone import BackboneSpec, KernelBackbone
    from oracle import medtsllm_oracle as O
    spec = BackboneSpec("llama", hidden=256, layers=2, heads=2, inter=384, vocab=128, eps=1e-5)
    bb = KernelBackbone.random_init(spec, cuda, seed=3)
    sd = {}
    for i, lay in enumerate(bb.layers):
        p = f"layers.{i}."
        wqkv = lay["wqkv"].float().cpu()
        sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.v_proj.weight"] = wqkv.split(256, 0)
        sd[p + "self_attn.o_proj.weight"] = lay["wo"].float().cpu()
        wgu = lay["wgu"].float().cpu().view(-1, 2, 128, 256)
        sd[p + "mlp.gate_proj.weight"] = wgu[:, 0].reshape(-1, 256)[:384]
        sd[p + "mlp.up_proj.weight"] = wgu[:, 1].reshape(-1, 256)[:384]
        sd[p + "mlp.down_proj.weight"] = lay["wdown"].float().cpu()[:, :384]
        sd[p + "input_layernorm.weight"] = lay["ln1"].cpu()
        sd[p + "post_attention_layernorm.weight"] = lay["ln2"].cpu()
    sd["norm.weight"] = bb.final_norm_w.cpu()
    Bp, L = 3, 70
    x = torch.randn(Bp, L, 256, generator=torch.Generator().manual_seed(0))
    ref = O.llama_forward(x, sd, n_layers=2, n_heads=2, eps=1e-5)
    got = bb.forward(x.to(cuda).view(Bp * L, 256).contiguous(), Bp, L)[0].float().cpu().view(Bp, L, 256)
    assert _rel_l2(got, ref) < 1e-2


@pytest.mark.parametrize("name", CASES)
def test_training_step_gradients(name, tmp_path, cuda):
    """loss.backward() through the kernel stack vs autograd through the oracle (fp32, CPU) on the same
    weights/inputs/loss.  Tolerance: relative L2 < 5e-2 per adapter tensor (bf16 operands in ~10
    chained GEMMs forward and backward; the largest tensors come out around 1e-2)."""
    from medtsllm_b200.model import MedTsLLM
    from oracle import medtsllm_oracle as O
    from _fixtures import oracle_spec
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    model = MedTsLLM(Cfg(config_for(fix, llm_dir)), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda, torch.float32).train()
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    out = model(inputs)
    assert out.requires_grad
    gen = torch.Generator().manual_seed(5)
    wgt = torch.randn(out.shape, generator=gen)
    (out * wgt.to(cuda)).sum().backward()
    torch.cuda.synchronize()

    # oracle gradients (prompt ids from the GPU-side host logic so both sides embed the same tokens)
    ids = prompt_items(model, inputs)
    ad = {k: v.clone().requires_grad_(True) for k, v in fix["adapters"].items()}
    sd = {k: v.float() for k, v in fix["backbone_state"].items()}
    ref = O.medtsllm_forward(fix["inputs"]["x_enc"], ids, ad, sd, oracle_spec(fix), training=True)
    assert _rel_l2(out, ref) < 2e-2
    (ref * wgt).sum().backward()
    report, bad = [], []
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        gref = ad[k].grad
        tag = f"{k.split('.')[-2][:8]}.{k.split('.')[-1][0]}"
        sib = k.rsplit(".", 1)[0] + ".weight"
        if k.endswith(".bias") and sib in ad and gref.abs().max() < 1e-5 * ad[sib].grad.abs().max():
            # Structurally ZERO gradients: d/d(key_projection.bias) (q.(k + b_k) shifts all scores of a query row
            # equally and softmax is shift invariant) and, with a GPT-2 backbone, d/d(feature_weighting.bias) (a
            # constant added to every feature of a token is removed by every LayerNorm that reads the residual
            # stream).  Both sides hold rounding noise only -> absolute check against the sibling weight gradient.
            scale = dict(model.named_parameters())[sib].grad.abs().max().item()
            e = p.grad.abs().max().item() / max(scale, 1e-30)
            report.append(f"{tag} |g|/|gW| {e:.1e}")
            if not e < 1e-1:
                bad.append((k, e))
            continue
        e = _rel_l2(p.grad, gref)
        report.append(f"{tag} {e:.1e}")
        # mapping_layer.bias: row sums with heavy cancellation, conditioning ~10x worse (see the LoRA test)
        if not e < (1.5e-1 if k == "mapping_layer.bias" else 5e-2):
            bad.append((k, e))
    print(f"\n[grad parity] {name}: " + "  ".join(report))
    assert not bad, (name, bad)
    # one optimizer step changes the cached bf16 weights (version tracking) and the output
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-2)
    opt.step()
    with torch.no_grad():
        out2 = model(inputs)
    assert not torch.equal(out2, out.detach())


@pytest.mark.parametrize("name,rank", [("llama_forecast_clip_stats", 8), ("gpt2_anomaly_concat", 4)])
def test_lora_forward_and_gradients(name, rank, tmp_path, cuda):
    """LoRA (BASELINE config 5).  peft is absent, so parity is pinned two ways: (i) identity at
    initialisation (B = 0) against the non-LoRA model, bit for bit; (ii) with random non-zero B, forward and
    all gradients (adapters + every A/B pair) against the oracle's plain-PyTorch LoRA restatement."""
    from medtsllm_b200.model import MedTsLLM
    from oracle import medtsllm_oracle as O
    from _fixtures import oracle_spec
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    base = MedTsLLM(Cfg(config_for(fix, llm_dir)), Dataset(fix["dataset"]))
    base.load_state_dict(fix["adapters"], strict=True)
    base = base.to(cuda).eval()
    cfg = config_for(fix, llm_dir)
    cfg["models"]["medtsllm"]["lora"] = {"enabled": True, "layers": "auto", "rank": rank, "alpha": 16, "rslora": True}
    torch.manual_seed(3)
    model = MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    assert model.lora_enabled and list(model.state_dict().keys()) == list(fix["adapters"].keys())
    res = model.load_state_dict(fix["adapters"], strict=False)          # as tasks/base.py:300 (LoRA pairs live outside)
    assert not res.unexpected_keys and all(k.startswith("llm.") for k in res.missing_keys)
    model = model.to(cuda).eval()
    with torch.no_grad():
        # (i) identity at init: B = 0 makes the LoRA update exactly zero.  Not bit-identical to `base` only because
        # the LoRA path rotates q/k with the separate RoPE kernel (on bf16 q/k, after the LoRA accumulate) while
        # the plain path rotates the fp32 accumulators in the QKV epilogue: rounding-level difference.
        assert _rel_l2(model(inputs), base(inputs)) < 2e-3
    with torch.no_grad():
        for b in model.llm.B:
            b.normal_(std=0.05)
    model.train()
    out = model(inputs)
    gen = torch.Generator().manual_seed(6)
    wgt = torch.randn(out.shape, generator=gen)
    (out * wgt.to(cuda)).sum().backward()
    ids = [row.tolist() for row in model.prompt_token_ids(inputs)]
    ad = {k: v.clone().requires_grad_(True) for k, v in fix["adapters"].items()}
    lora = {"scale": model.llm.scale, "n_targets": len(model.llm.targets),
            "A": [a.detach().cpu().clone().requires_grad_(True) for a in model.llm.A],
            "B": [b.detach().cpu().clone().requires_grad_(True) for b in model.llm.B]}
    sd = {k: v.float() for k, v in fix["backbone_state"].items()}
    ref = O.medtsllm_forward(fix["inputs"]["x_enc"], ids, ad, sd, oracle_spec(fix), training=True, lora=lora)
    assert _rel_l2(out, ref) < 2e-2
    assert _rel_l2(out, base.train()(inputs).detach()) > 1e-3               # the adapters do change the output
    (ref * wgt).sum().backward()
    errs = {}
    for k, p in model.named_parameters():
        if k.startswith("llm.") or k == "reprogramming_layer.key_projection.bias":
            continue
        errs[k] = _rel_l2(p.grad, ad[k].grad)
    for i, (a, b) in enumerate(zip(model.llm.A, model.llm.B)):
        errs[f"lora_A[{i}]"] = _rel_l2(a.grad, lora["A"][i].grad)
        errs[f"lora_B[{i}]"] = _rel_l2(b.grad, lora["B"][i].grad)
    # d b_map[s] = sum_d dSource[s, d] is a sum with heavy cancellation (|sum| << sum|.|): its conditioning is
    # ~10x worse than every other tensor's, so bf16 noise upstream shows up amplified -> its own tolerance
    tol = {k: 5e-2 for k in errs}
    tol["mapping_layer.bias"] = 1.5e-1
    worst = max(errs.items(), key=lambda kv: kv[1] / tol[kv[0]])
    print(f"\n[lora parity] {name}: worst {worst[0]} {worst[1]:.1e}; lora_A[0] {errs['lora_A[0]']:.1e} lora_B[0] {errs['lora_B[0]']:.1e}")
    assert worst[1] < tol[worst[0]], worst
    model.llm.save_pretrained(tmp_path / "ckpt" / "best-lora.safetensors")
    assert (tmp_path / "ckpt" / "best-lora.safetensors").exists()


def test_training_with_dropout(tmp_path, cuda):
    """training.dropout = 0.1 (the shipped configs): the masks the kernel path draws are read back (same counter-based
    function of the step's seeds) and handed to the oracle, so forward and gradients can be compared exactly like the
    deterministic case; plus the drop rate itself."""
    from medtsllm_b200 import ops
    from medtsllm_b200.model import MedTsLLM
    from oracle import medtsllm_oracle as O
    from _fixtures import oracle_spec
    fix = load_case("llama_seg_concat")
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    cfg = config_for(fix, llm_dir)
    cfg["training"]["dropout"] = 0.1
    model = MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda).train()
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    torch.manual_seed(11)
    out = model(inputs)
    s_patch, s_attn = model._last_dropout_seeds
    wgt = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * wgt.to(cuda)).sum().backward()
    B, T, C = fix["inputs"]["x_enc"].shape
    N, H, S, p = model.n_patches, model.n_attention_heads, model.num_tokens, 0.1
    keep_patch = ops.dropout(torch.ones(B, N, C * 32, device=cuda, dtype=torch.bfloat16), p, s_patch) != 0
    keep_attn = ops.dropout(torch.ones(H, B * N, S, device=cuda, dtype=torch.bfloat16), p, s_attn) != 0
    assert abs(1 - keep_attn.float().mean().item() - p) < 2e-3 and abs(1 - keep_patch.float().mean().item() - p) < 3e-2
    # oracle layouts: patch mask in [B*C, N, 32] (pre-concat) order, attention mask [B, H, N, S]
    m_patch = (keep_patch.float().cpu() / (1 - p)).view(B, N, C, 32).permute(0, 2, 1, 3).reshape(B * C, N, 32)
    m_attn = (keep_attn.float().cpu() / (1 - p)).view(H, B, N, S).permute(1, 0, 2, 3)
    ids = [row.tolist() for row in model.prompt_token_ids(inputs)]
    ad = {k: v.clone().requires_grad_(True) for k, v in fix["adapters"].items()}
    sd = {k: v.float() for k, v in fix["backbone_state"].items()}
    ref = O.medtsllm_forward(fix["inputs"]["x_enc"], ids, ad, sd, oracle_spec(fix), training=True,
                             dropout_masks={"patch": m_patch, "reprog": m_attn})
    assert _rel_l2(out, ref) < 2e-2
    assert _rel_l2(out, fix["stages"]["output_train"]) > 1e-2        # dropout really changed the forward
    (ref * wgt).sum().backward()
    for k, prm in model.named_parameters():
        if k in ("reprogramming_layer.key_projection.bias",):
            continue
        e = _rel_l2(prm.grad, ad[k].grad)
        assert e < (1.5e-1 if k == "mapping_layer.bias" else 5e-2), (k, e)
    # eval mode: no dropout, deterministic
    model.eval()
    with torch.no_grad():
        assert torch.equal(model(inputs), model(inputs)) and model._last_dropout_seeds is None


@pytest.mark.parametrize("name", ["gpt2_anomaly_concat", "llama_seg_concat"])
def test_training_with_backbone_dropout(name, tmp_path, cuda):
    """The frozen backbone's own dropouts (GPT-2 checkpoints: embd / attn / resid 0.1, live in the reference's train mode
    because tasks/forecasting.py:18 flips the HF module; Llama: attention_dropout) on the kernel path: the counter-based
    masks are read back from the step's seeds and replayed in the oracle (whose placement of them is pinned against live
    HF, tests/test_oracle.py), so forward and adapter gradients compare like the deterministic case."""
    from medtsllm_b200 import ops
    from medtsllm_b200.model import MedTsLLM
    from oracle import medtsllm_oracle as O
    from _fixtures import oracle_spec
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    model = MedTsLLM(Cfg(config_for(fix, llm_dir)), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda).train()
    p = 0.1
    gpt2 = fix["kind"] == "gpt2"
    model.backbone_dropout = {"embd": p, "attn": p, "resid": p} if gpt2 else {"embd": 0.0, "attn": p, "resid": 0.0}
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    torch.manual_seed(3)
    out = model(inputs)
    drop = model._last_backbone_dropout
    wgt = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * wgt.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    spec = oracle_spec(fix)
    ids = [row.tolist() for row in model.prompt_token_ids(inputs)]
    B = fix["inputs"]["x_enc"].shape[0]
    L = len(ids[0]) + model.n_patches
    D, H, n_layers = model.d_llm, spec["llm_heads"], spec["n_layers"]
    seeds = drop["seeds"]
    assert len(seeds) == 1 + 3 * n_layers

    def mask(shape, prob, seed):
        if prob <= 0:
            return torch.ones(shape)
        keep = ops.dropout(torch.ones(shape, device=cuda), prob, seed) != 0
        assert abs(1 - keep.float().mean().item() - prob) < 2e-2
        return keep.float().cpu() / (1 - prob)

    masks = {"attn": [mask((B, H, L, L), p, seeds[1 + 3 * i]) for i in range(n_layers)]}
    if gpt2:
        masks.update(embd=mask((B, L, D), p, seeds[0]),
                     resid_attn=[mask((B, L, D), p, seeds[2 + 3 * i]) for i in range(n_layers)],
                     resid_mlp=[mask((B, L, D), p, seeds[3 + 3 * i]) for i in range(n_layers)])
    ad = {k: v.clone().requires_grad_(True) for k, v in fix["adapters"].items()}
    sd = {k: v.float() for k, v in fix["backbone_state"].items()}
    ref = O.medtsllm_forward(fix["inputs"]["x_enc"], ids, ad, sd, spec, training=True, dropout_masks={"backbone": masks})
    e = _rel_l2(out, ref)
    assert e < 2e-2, e
    assert _rel_l2(out, fix["stages"]["output_train"]) > 5e-3          # the dropouts really changed the forward
    (ref * wgt).sum().backward()
    for k, prm in model.named_parameters():
        if k in ("reprogramming_layer.key_projection.bias",):
            continue
        eg = _rel_l2(prm.grad, ad[k].grad)
        assert eg < (1.5e-1 if k == "mapping_layer.bias" else 5e-2), (k, eg)
    # evaluation: no dropout, shared prompt prefix and graph replay are back
    model.eval()
    with torch.no_grad():
        assert torch.equal(model(inputs), model(inputs)) and model._last_backbone_dropout is None


def test_edge_inputs_empty_prompt_2d_input_batch_of_one(tmp_path, cuda):
    """Edge cases of the reference path: no prompt at all (models/medtsllm.py:338-339 -> [B, 0, D]), a 2-D univariate
    window tensor (:264-265), and a batch of one."""
    from medtsllm_b200.model import MedTsLLM
    from oracle import medtsllm_oracle as O
    from _fixtures import oracle_spec
    fix = load_case("llama_semseg_univariate")
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    cfg = config_for(fix, llm_dir)
    cfg["models"]["medtsllm"]["prompting"].update(dataset=False, task=False, clip=False, input_stats=False)
    model = MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda).eval()
    x = fix["inputs"]["x_enc"][:1]                                   # [1, T, 1]
    sd = {k: v.float() for k, v in fix["backbone_state"].items()}
    ref = O.medtsllm_forward(x, [[]], fix["adapters"], sd, oracle_spec(fix))
    with torch.no_grad():
        out3 = model({"x_enc": x.to(cuda)})
        out2 = model({"x_enc": x[:, :, 0].to(cuda)})                 # 2-D input -> unsqueeze
    assert out3.shape == ref.shape and torch.equal(out2, out3)
    assert _rel_l2(out3, ref) < 2e-2
    # and it trains: one backward with an empty prompt
    model.train()
    out = model({"x_enc": x.to(cuda)})
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_llm_layers_truncates_the_backbone(tmp_path, cuda):
    """`models.medtsllm.llm.llm_layers = k` keeps the first k blocks of the checkpoint (models/medtsllm.py:144-146:
    `llm_config.num_hidden_layers = llm_layers` before `AutoModel.from_pretrained`)."""
    from medtsllm_b200.model import MedTsLLM
    from _fixtures import oracle_spec
    fix = load_case("llama_seg_concat")                                  # 2-block backbone
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    cfg = config_for(fix, llm_dir)
    cfg["models"]["medtsllm"]["llm"]["llm_layers"] = 1
    model = MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda).eval()
    assert len(model._backbone.layers) == 1
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    with torch.no_grad():
        out = model(inputs)
    from oracle import medtsllm_oracle as O
    spec = dict(oracle_spec(fix), n_layers=1)
    ids = [row.tolist() for row in model.prompt_token_ids(inputs)]
    ref = O.medtsllm_forward(fix["inputs"]["x_enc"], ids, fix["adapters"], {k: v.float() for k, v in fix["backbone_state"].items()}, spec)
    full = run_oracle(fix)[0]
    assert _rel_l2(out, ref) < 3e-3
    assert _rel_l2(ref, full) > 1e-2                                    # the second block does matter


def test_long_sequence_falls_back_to_tiled_attention(cuda):
    """L beyond the shared-memory-resident attention kernels (hd 128: L > ~350) takes the 64x64-tiled kernels
    (forward and backward) and still matches the oracle's Llama restatement."""
    from medtsllm_b200.backbone import BackboneSpec, KernelBackbone
    from oracle import medtsllm_oracle as O
    spec = BackboneSpec("llama", hidden=256, layers=1, heads=2, inter=256, vocab=64, eps=1e-5, max_pos=1024)
    bb = KernelBackbone.random_init(spec, cuda, seed=5)
    lay = bb.layers[0]
    wqkv = lay["wqkv"].float().cpu()
    wgu = lay["wgu"].float().cpu().view(-1, 2, 128, 256)
    sd = {"layers.0.self_attn.q_proj.weight": wqkv[:256], "layers.0.self_attn.k_proj.weight": wqkv[256:512],
          "layers.0.self_attn.v_proj.weight": wqkv[512:], "layers.0.self_attn.o_proj.weight": lay["wo"].float().cpu(),
          "layers.0.mlp.gate_proj.weight": wgu[:, 0].reshape(-1, 256)[:256],
          "layers.0.mlp.up_proj.weight": wgu[:, 1].reshape(-1, 256)[:256],
          "layers.0.mlp.down_proj.weight": lay["wdown"].float().cpu()[:, :256],
          "layers.0.input_layernorm.weight": lay["ln1"].cpu(), "layers.0.post_attention_layernorm.weight": lay["ln2"].cpu(),
          "norm.weight": bb.final_norm_w.cpu()}
    Bp, L = 2, 420
    x = torch.randn(Bp, L, 256, generator=torch.Generator().manual_seed(1))
    xr = x.clone().requires_grad_(True)
    ref = O.llama_forward(xr, sd, n_layers=1, n_heads=2, eps=1e-5)
    stash = []
    got, x_final = bb.forward(x.to(cuda).view(Bp * L, 256).contiguous(), Bp, L, stash=stash)
    assert _rel_l2(got.float().cpu().view(Bp, L, 256), ref) < 1e-2
    w = torch.randn(Bp, L, 256, generator=torch.Generator().manual_seed(2))
    (ref * w).sum().backward()
    dR, _ = bb.backward(w.to(cuda, torch.bfloat16).view(Bp * L, 256).contiguous(), x_final, stash, Bp, L)
    assert _rel_l2(dR.cpu().view(Bp, L, 256), xr.grad) < 2e-2


@pytest.mark.parametrize("name,partial", [("llama_seg_concat", 0), ("gpt2_anomaly_concat", 0), ("llama_forecast_independent", 0),
                                          ("llama_forecast_interleave", 0), ("gpt2_anomaly_weighted_average", 0),
                                          ("llama_seg_concat", 21), ("gpt2_forecast_merge_end", 17)])
def test_shared_prompt_prefix_equals_per_sample_prompts(name, partial, tmp_path, cuda):
    """In-batch prompt de-duplication (shared-prefix row layout): the leading prompt positions that are the
    same in every sample are carried once through the backbone.  Output and every adapter gradient must equal
    the plain per-sample computation (the reference's, models/medtsllm.py:330-351) — the kernels do the same
    arithmetic per row in the same order, so the forward is compared bit for bit.  `partial` > 0 makes the
    prompts differ from that position on (per-sample clip descriptions / input statistics)."""
    from medtsllm_b200.model import MedTsLLM
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    model = MedTsLLM(Cfg(config_for(fix, llm_dir)), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda, torch.float32)
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    if partial:
        table = model.prompt_token_ids(inputs).clone()
        g = torch.Generator().manual_seed(partial)
        table[:, partial:] = torch.randint(3, model.vocab_size, table[:, partial:].shape, generator=g, dtype=torch.int32)
        table[1, partial] = table[0, partial] + 1          # the prefix ends exactly here
        model.prompt_token_ids = lambda _inputs: table
    res = {}
    for share in (True, False):
        model.share_prompt_prefix = share
        model.eval()
        model._capture = {}
        with torch.no_grad():
            out = model(inputs)
        cap, model._capture = model._capture, None
        model.train()
        model.zero_grad(set_to_none=True)
        y = model(inputs)
        wgt = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(cuda)
        (y * wgt).sum().backward()
        torch.cuda.synchronize()
        res[share] = (out, cap, {k: p.grad.clone() for k, p in model.named_parameters()})
    (o1, c1, g1), (o0, c0, g0) = res[True], res[False]
    assert c0["shared_prefix"] == 0
    assert c1["shared_prefix"] == (partial if partial else len(fix["prompt_ids"][0])), c1["shared_prefix"]
    assert torch.equal(c1["llm_input"], c0["llm_input"])
    assert torch.equal(c1["llm"], c0["llm"])
    assert torch.equal(o1, o0)
    for k in g0:
        # the backward skips the prefix rows (no trainable ancestor): same sums, fewer zero terms
        e = _rel_l2(g1[k], g0[k])
        # structurally zero gradients (key-projection bias, ...: see test_training_step_gradients) hold rounding
        # noise on both sides -> absolute check against the sibling weight gradient
        sib = g0.get(k.rsplit(".", 1)[0] + ".weight", g0[k]).abs().max().item()
        assert e < 2e-3 or (g1[k] - g0[k]).abs().max().item() <= 1e-3 * sib, (k, e)


@pytest.mark.parametrize("name,rank", [("llama_seg_concat", 8), ("gpt2_anomaly_concat", 4)])
def test_lora_training_on_shared_prefix_equals_per_sample_prompts(name, rank, tmp_path, cuda):
    """LoRA fine-tuning (BASELINE config 5) keeps the shared-prefix layout: the A/B pairs receive gradient through the
    prompt rows too, so the backward runs on all rows and the prefix keys collect dK / dV from every sample
    (attn_bwd_dkv_prefix_kernel).  Every gradient — adapters and all LoRA pairs — must equal the per-sample-prompt run."""
    from medtsllm_b200.model import MedTsLLM
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    cfg = config_for(fix, llm_dir)
    cfg["models"]["medtsllm"]["lora"] = {"enabled": True, "layers": "auto", "rank": rank, "alpha": 16, "rslora": True}
    torch.manual_seed(3)
    model = MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=False)
    model = model.to(cuda, torch.float32).train()
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in model.llm.B:                      # non-zero B: the LoRA path is live
            p.copy_((torch.randn(p.shape, generator=gen) * 0.05).to(cuda))
    inputs = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    res = {}
    for share in (True, False):
        model.share_prompt_prefix = share
        model.zero_grad(set_to_none=True)
        y = model(inputs)
        wgt = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(cuda)
        (y * wgt).sum().backward()
        torch.cuda.synchronize()
        named = dict(model.named_parameters())
        res[share] = (y.detach().clone(), {k: p.grad.clone() for k, p in named.items() if p.grad is not None})
    (y1, g1), (y0, g0) = res[True], res[False]
    assert torch.equal(y1, y0)
    assert set(g1) == set(g0) and any(k.startswith("llm.") for k in g1)
    worst = {}
    for k in g0:
        e = _rel_l2(g1[k], g0[k])
        sib = g0.get(k.rsplit(".", 1)[0] + ".weight", g0[k]).abs().max().item()
        worst[k] = e
        # per-sample prompts round every copy's prefix contribution to bf16 before summing; shared rows sum in fp32
        assert e < 2e-2 or (g1[k] - g0[k]).abs().max().item() <= 1e-3 * sib, (k, e)
    print(f"\n[lora shared-prefix] {name}: worst rel-L2 {max(worst.values()):.2e} over {len(worst)} tensors")


@pytest.mark.parametrize("name", ["llama_seg_concat", "gpt2_anomaly_concat", "llama_forecast_clip_stats"])
def test_cuda_graph_replay_matches_kernel_by_kernel(name, tmp_path, cuda):
    """Inference replays a captured CUDA graph once the same (shape, prompt table, weights) key repeats.  Call 1 runs
    kernel by kernel, call 2 captures, call 3 replays; each must equal the kernel-by-kernel result for ITS input bit
    for bit, and an optimizer-style in-place weight update must invalidate the graph.  (clip_stats: prompts change
    with every batch -> never captured.)"""
    from medtsllm_b200 import _lib
    from medtsllm_b200.model import MedTsLLM
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    model = MedTsLLM(Cfg(config_for(fix, llm_dir)), Dataset(fix["dataset"]))
    model.load_state_dict(fix["adapters"], strict=True)
    model = model.to(cuda, torch.float32).eval()
    base = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}
    xs = [base["x_enc"] * (1.0 + 0.1 * i) + 0.3 * i for i in range(4)]

    def run(x, graph):
        model.use_cuda_graph = graph
        with torch.no_grad():
            return model({**base, "x_enc": x}).clone()

    want = [run(x, False) for x in xs]
    n0 = _lib.launch_count()
    got = [run(x, True) for x in xs]
    per_call = (_lib.launch_count() - n0) / len(xs)
    for w, g in zip(want, got):
        assert torch.equal(w, g)
    dynamic = name == "llama_forecast_clip_stats"
    assert model._graph.captured == (not dynamic)
    assert per_call > 20                                        # replays are counted as launches too
    with torch.no_grad():
        model.output_projection.linear.bias.add_(0.25)          # what optimizer.step() does: bumps ._version
    after = run(xs[3], True)
    assert torch.equal(after, run(xs[3], False)) and not torch.equal(after, got[3])


# ------------------------------------------------------------------------------------------------ GPT4TS
from _fixtures import GPT4TS_CASES, gpt4ts_hf_model, gpt4ts_spec, load_gpt4ts_backbone  # noqa: E402


@pytest.mark.parametrize("name", GPT4TS_CASES)
def test_gpt4ts_forward_parity(name, cuda):
    """medtsllm_b200.GPT4TS (BASELINE configs[0] = gpt4ts_forecast_etth1) against the golden outputs of the
    unmodified models/gpt4ts.py and against the oracle.  Tolerance: relative L2 < 5e-3 on the output (bf16 GEMM
    operands with fp32 accumulation through two GPT-2 blocks, fp32 residual stream; both references are fp32)."""
    from medtsllm_b200._lib import MtsError
    from medtsllm_b200.backbone import KernelBackbone
    from medtsllm_b200.gpt4ts import GPT4TS
    from oracle import gpt4ts_oracle as G
    fix = load_case(name)
    bbf = load_gpt4ts_backbone()
    n_layers = fix["config"]["models"]["gpt4ts"]["gpt_layers"]
    backbone = KernelBackbone.from_hf(gpt4ts_hf_model(bbf, n_layers), cuda)
    model = GPT4TS(Cfg(fix["config"]), Dataset(fix["dataset"]), backbone=backbone)
    res = model.load_state_dict(fix["params"], strict=False)
    assert not res.unexpected_keys and all(k == "enc_embedding.position_embedding.pe" or k.startswith("gpt2.")
                                           for k in res.missing_keys), res
    model = model.to(cuda, torch.float32).eval()
    x = fix["inputs"]["x_enc"].to(cuda)
    with torch.no_grad():
        out = model({"x_enc": x.clone()})
        model.train()
        out_t = model({"x_enc": x.clone()})
        model.eval()
    g = fix["stages"]
    ref = G.gpt4ts_forward(fix["inputs"]["x_enc"], fix["params"], {k: v.float() for k, v in bbf["state"].items()},
                           gpt4ts_spec(fix, bbf))
    assert out.shape == g["output"].shape and out.dtype == torch.float32
    e_g, e_o, e_t = _rel_l2(out, g["output"]), _rel_l2(out, ref), _rel_l2(out_t, g["output_train"])
    print(f"\n[gpt4ts parity] {name}: rel-L2 vs golden {e_g:.2e}, vs oracle {e_o:.2e}, train-mode {e_t:.2e}")
    assert e_g < 5e-3 and e_o < 5e-3 and e_t < 5e-3, (name, e_g, e_o, e_t)
    if fix["config"]["task"] == "anomaly_detection":
        # the prediction is x + sqrt(1e-5) * dec: check the model part on its own as well
        xin = fix["inputs"]["x_enc"]
        assert _rel_l2(out.cpu() - xin, g["output"] - xin) < 1e-2
    with pytest.raises(MtsError):
        with torch.no_grad():
            model({"x_enc": x.cpu()})


@pytest.mark.parametrize("name", GPT4TS_CASES)
def test_gpt4ts_training_gradients(name, cuda):
    """loss.backward() through medtsllm_b200.GPT4TS vs autograd through the oracle (fp32, CPU): the model's own tensors and
    the GPT-2 tensors the reference trains (every LayerNorm weight / bias and the position table, models/gpt4ts.py:47-53).
    Tolerance: relative L2 < 5e-2 per tensor (bf16 operands forward and backward); one Adam step must move the
    LayerNorms inside the kernel backbone (the parameters share its storage)."""
    from medtsllm_b200.backbone import KernelBackbone
    from medtsllm_b200.gpt4ts import GPT4TS
    from oracle import gpt4ts_oracle as G
    fix = load_case(name)
    bbf = load_gpt4ts_backbone()
    n_layers = fix["config"]["models"]["gpt4ts"]["gpt_layers"]
    backbone = KernelBackbone.from_hf(gpt4ts_hf_model(bbf, n_layers), cuda)
    model = GPT4TS(Cfg(fix["config"]), Dataset(fix["dataset"]), backbone=backbone)
    model.load_state_dict(fix["params"], strict=False)
    model = model.to(cuda, torch.float32).train()
    x = fix["inputs"]["x_enc"]
    out = model({"x_enc": x.to(cuda)})
    assert out.requires_grad
    wgt = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * wgt.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    # oracle side
    params = {k: v.clone().requires_grad_(True) for k, v in fix["params"].items()}
    sd = {k: v.float().requires_grad_(k.startswith("ln_f") or ".ln_" in k or k.startswith("wpe")) for k, v in bbf["state"].items()}
    ref = G.gpt4ts_forward(x, params, sd, gpt4ts_spec(fix, bbf), training=True)
    assert _rel_l2(out, ref) < 1e-2
    (ref * wgt).sum().backward()
    report, bad, checked = [], [], 0
    for k, p in model.named_parameters():
        gref = sd[k[5:]].grad if k.startswith("gpt2.") else params[k].grad
        if gref is None or gref.abs().max() == 0:
            assert p.grad is None or p.grad.abs().max().item() == 0, k        # unused by the forward on both sides
            continue
        assert p.grad is not None, k
        checked += 1
        tag = k.replace('gpt2.', '').replace('enc_embedding.value_embedding.', '')
        sib = k.rsplit(".", 1)[0] + ".weight"
        if k == "predict_linear_pre.bias" and gref.abs().max() < 1e-4 * params[sib].grad.abs().max():
            # structurally ZERO: the bias shifts all d_model features of a token by the same amount, which every LayerNorm
            # reading the residual stream (ln_1, ln_2, ln_f) removes.  Both sides hold rounding noise only.
            scale = dict(model.named_parameters())[sib].grad.abs().max().item()
            e = p.grad.abs().max().item() / max(scale, 1e-30)
            report.append(f"{tag} |g|/|gW| {e:.1e}")
            if not e < 1e-1:
                bad.append((k, e))
            continue
        e = _rel_l2(p.grad, gref)
        report.append(f"{tag} {e:.1e}")
        if not e < 5e-2:
            bad.append((k, e))
    print(f"\n[gpt4ts grad parity] {name}: " + "  ".join(report))
    assert not bad and checked >= 5 + 4 * n_layers, (name, bad, checked)
    before = backbone.layers[0]["ln1"].clone()
    torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-2).step()
    assert not torch.equal(backbone.layers[0]["ln1"], before)


@pytest.mark.parametrize("name", ["gpt4ts_forecast_etth1", "gpt4ts_semseg", "gpt4ts_anomaly"])
def test_gpt4ts_training_with_dropout(name, cuda):
    """the reference's GPT4TS trains with DataEmbedding.dropout(training.dropout) and with GPT-2's own 0.1
    embd / attn / resid dropouts (model.train() flips the HF module).  Same scheme as MedTsLLM: counter-based masks read
    back from the step's seeds and replayed in the oracle; forward and every trained tensor's gradient are compared."""
    import copy
    from medtsllm_b200 import ops
    from medtsllm_b200.backbone import KernelBackbone
    from medtsllm_b200.gpt4ts import GPT4TS
    from oracle import gpt4ts_oracle as G
    fix = load_case(name)
    bbf = load_gpt4ts_backbone()
    cfg = copy.deepcopy(fix["config"])
    p_emb, p_bb = 0.1, 0.1
    cfg["training"]["dropout"] = p_emb
    n_layers = cfg["models"]["gpt4ts"]["gpt_layers"]
    backbone = KernelBackbone.from_hf(gpt4ts_hf_model(bbf, n_layers), cuda)
    model = GPT4TS(Cfg(cfg), Dataset(fix["dataset"]), backbone=backbone)
    model.load_state_dict(fix["params"], strict=False)
    model = model.to(cuda, torch.float32).train()
    model.backbone_dropout = {"embd": p_bb, "attn": p_bb, "resid": p_bb}
    x = fix["inputs"]["x_enc"]
    torch.manual_seed(7)
    out = model({"x_enc": x.to(cuda)})
    drop = model._last_dropout
    wgt = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * wgt.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    spec = gpt4ts_spec(fix, bbf)
    B, T, C = x.shape
    D, H = backbone.spec.hidden, backbone.spec.heads
    T2 = T + (cfg["pred_len"] if cfg["task"] == "forecasting" else 0)
    dm = cfg["models"]["gpt4ts"]["d_model"]

    def mask(shape, prob, seed, dtype=torch.float32):
        keep = ops.dropout(torch.ones(shape, device=cuda, dtype=dtype), prob, seed) != 0
        return keep.float().cpu() / (1 - prob)

    seeds = drop["bb"]["seeds"]
    bbm = {"embd": mask((B, T2, D), p_bb, seeds[0]),
           "attn": [mask((B, H, T2, T2), p_bb, seeds[1 + 3 * i]) for i in range(n_layers)],
           "resid_attn": [mask((B, T2, D), p_bb, seeds[2 + 3 * i]) for i in range(n_layers)],
           "resid_mlp": [mask((B, T2, D), p_bb, seeds[3 + 3 * i]) for i in range(n_layers)]}
    emb_mask = None
    if cfg["task"] != "anomaly_detection":
        # forecasting: dropped on the bf16 [B, T, d_model] copy; segmentation tasks: on the fp32 [B, T, D] rows (d_model = D)
        emb_mask = mask((B, T, dm), p_emb, drop["seed_embed"], torch.bfloat16 if cfg["task"] == "forecasting" else torch.float32)
    params = {k: v.clone().requires_grad_(True) for k, v in fix["params"].items()}
    sd = {k: v.float().requires_grad_(k.startswith("ln_f") or ".ln_" in k or k.startswith("wpe")) for k, v in bbf["state"].items()}
    ref = G.gpt4ts_forward(x, params, sd, spec, training=True, dropout_masks={"embed": emb_mask, "backbone": bbm})
    assert _rel_l2(out, ref) < 2e-2, _rel_l2(out, ref)
    (ref * wgt).sum().backward()
    checked = 0
    for k, p in model.named_parameters():
        gref = sd[k[5:]].grad if k.startswith("gpt2.") else params[k].grad
        if gref is None or gref.abs().max() == 0 or p.grad is None:
            continue
        if k == "predict_linear_pre.bias":            # structurally zero (see test_gpt4ts_training_gradients)
            continue
        e = _rel_l2(p.grad, gref)
        assert e < 6e-2, (k, e)
        checked += 1
    assert checked >= 4 + 4 * n_layers
    model.eval()
    with torch.no_grad():
        assert torch.equal(model({"x_enc": x.to(cuda)}), model({"x_enc": x.to(cuda)})) and model._last_dropout is None


@pytest.mark.parametrize("name,lora", [("llama_seg_concat", False), ("gpt2_anomaly_concat", False), ("llama_seg_concat", True)])
def test_training_step_graphs_match_kernel_by_kernel(name, lora, tmp_path, cuda):
    """Training steps replay two captured CUDA graphs (forward + stash, backward chain) once the same step shape
    repeats: step 0 runs kernel by kernel, step 1 captures, steps 2.. replay.  Losses and the parameters after
    5 SGD steps must equal the kernel-by-kernel run (same kernels, same order), the captured forward
    must pick up every optimizer update (bf16 re-casts are part of the graph), and a backward() whose activations
    were overwritten by a later forward must raise."""
    from medtsllm_b200._lib import MtsError
    from medtsllm_b200.model import MedTsLLM
    fix = load_case(name)
    llm_dir = materialize_llm_dir(fix, tmp_path / "llm")
    cfg = config_for(fix, llm_dir)
    if lora:
        cfg["models"]["medtsllm"]["lora"] = {"enabled": True, "layers": "auto", "rank": 8, "alpha": 16, "rslora": True}
    base = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in fix["inputs"].items()}

    def run(mode):
        torch.manual_seed(3)
        model = MedTsLLM(Cfg(cfg), Dataset(fix["dataset"]))
        model.load_state_dict(fix["adapters"], strict=False)
        model = model.to(cuda, torch.float32).train()
        model.use_train_graph = mode
        if lora:
            gen = torch.Generator().manual_seed(11)
            with torch.no_grad():
                for p in model.llm.B:
                    p.copy_((torch.randn(p.shape, generator=gen) * 0.05).to(cuda))
        # plain SGD: the update is linear in the gradient, so the last-bit run-to-run noise of the atomically reduced conv
        # gradient stays last-bit (Adam's g / sqrt(v) turns noise-only gradients — the structurally zero key bias — into
        # +-lr steps and makes any comparison of the two runs meaningless)
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-2)
        losses = []
        for step in range(5):
            y = model({**base, "x_enc": base["x_enc"] * (1.0 + 0.05 * step)})
            loss = y.float().pow(2).mean()
            loss.backward()
            opt.step()
            opt.zero_grad()
            losses.append(loss.item())
        return model, losses

    m0, l0 = run("0")
    m1, l1 = run("1")
    assert m0._train_graph is None and m1._train_graph.entry is not None and m1._train_graph.entry["bwd"] is not None
    assert all(abs(a - b) <= 1e-4 * abs(a) for a, b in zip(l0, l1)), (l0, l1)
    assert l0[0] != l0[-1]
    for (k, p0), (_, p1) in zip(m0.named_parameters(), m1.named_parameters()):
        # (not torch.equal: the patch-embedding conv gradient is reduced with fp32 atomics, whose summation order —
        # hence the last bits of that gradient and of everything Adam derives from it — varies from run to run)
        torch.testing.assert_close(p0, p1, rtol=1e-4, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")
    # evaluation after training on the graph sees the trained weights
    m0.eval(); m1.eval()
    with torch.no_grad():
        torch.testing.assert_close(m0(base), m1(base), rtol=1e-3, atol=1e-5)
    # stale activations: two forwards on the graph, backward through the first
    m1.train()
    ya = m1(base)
    m1(base)
    with pytest.raises(MtsError):
        ya.sum().backward()


def test_grouped_query_backbone_matches_live_hf(cuda):
    """Grouped-query checkpoints (fewer K / V heads than query heads): from_hf folds HuggingFace's repeat_kv
    (HF:models/llama/modeling_llama.py:186-196) into the K / V projection weights; the kernels run unchanged.  Against the
    live HuggingFace model in fp32 on the same GPU."""
    from transformers import LlamaConfig, LlamaModel
    from medtsllm_b200.backbone import KernelBackbone
    torch.manual_seed(11)
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=128, rms_norm_eps=1e-5, max_position_embeddings=256, attn_implementation="eager")
    hf = LlamaModel(cfg).eval()
    bb = KernelBackbone.from_hf(hf, cuda)
    assert bb.spec.kv_heads == 2 and bb.spec.heads == 4
    Bp, L = 3, 40
    x = torch.randn(Bp, L, 256) * 0.5
    with torch.no_grad():
        ref = hf.to(cuda)(inputs_embeds=x.to(cuda)).last_hidden_state
    out, _ = bb.forward(x.view(Bp * L, 256).to(cuda).clone(), Bp, L)
    err = ((out.float().view(Bp, L, 256) - ref).norm() / ref.norm()).item()
    assert err < 2e-2, err                                  # bf16 operands, two layers
