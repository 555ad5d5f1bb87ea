"""Pins the oracle (oracle/medtsllm_oracle.py, the CPU restatement that travels to the GPU box)
against (a) the golden stage tensors captured from the UNMODIFIED reference (tests/golden/*.pt, made
by oracle/make_golden.py) and (b) HuggingFace's own backbones executed here.  CPU only.

Tolerances: both sides are fp32 on CPU; differences come from summation order only
(conv-as-3-matmuls vs cuDNN-free Conv1d, fused vs split projections): rtol 1e-4 / atol 2e-5 on stage
tensors, with the final outputs additionally bounded in relative L2 (< 2e-5).
"""
import math

import pytest
import torch

from _fixtures import CASES, hf_model_from_fixture, load_case, oracle_spec, run_oracle
from oracle import medtsllm_oracle as O


def _rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    fix = load_case(name)
    out, st = run_oracle(fix)
    g = fix["stages"]
    B = fix["inputs"]["x_enc"].shape[0]
    checks = {
        "revin_mean": (st["revin_mean"], g["revin_mean"]),
        "revin_stdev": (st["revin_stdev"], g["revin_stdev"]),
        "patch_embedding": (st["patch_embedding"], g["patch_embedding"]),
        "source_embeddings": (st["source_embeddings"], g["source_embeddings"]),
        "reprogramming_layer": (st["reprogramming_layer"], g["reprogramming_layer"]),
        "llm_input": (st["llm_input"], g["llm_input"]),
        "llm": (st["llm"], g["llm"]),
        **({"downsample": (st["downsample"], g["downsample"])} if g.get("downsample") is not None else {}),
        "output_projection": (st["output_projection"], g["output_projection"]),
        "output": (out, g["output"]),
    }
    for key, (a, b) in checks.items():
        assert a.shape == b.shape, key
        torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-5, msg=lambda m, key=key: f"{name}/{key}: {m}")
    # per-layer hidden states (HF hidden_states = input + every block; the last entry is post-final-norm)
    hs_o, hs_g = st["llm.hidden_states"], g["llm.hidden_states"]
    assert len(hs_o) == len(hs_g)
    for i in range(len(hs_o) - 1):
        torch.testing.assert_close(hs_o[i], hs_g[i], rtol=1e-4, atol=2e-5)
    assert _rel_l2(out, g["output"]) < 2e-5
    out_t, _ = run_oracle(fix, training=True)
    torch.testing.assert_close(out_t, g["output_train"], rtol=1e-4, atol=2e-5)
    assert out.shape[0] == B


@pytest.mark.parametrize("name", ["llama_seg_concat", "gpt2_anomaly_concat"])
def test_oracle_backbone_matches_live_hf(name):
    fix = load_case(name)
    hf = hf_model_from_fixture(fix)
    spec = oracle_spec(fix)
    sd = {k: v.float() for k, v in fix["backbone_state"].items()}
    x = fix["stages"]["llm_input"]
    with torch.no_grad():
        ref = hf(inputs_embeds=x).last_hidden_state
    fwd = O.llama_forward if fix["kind"] == "llama" else O.gpt2_forward
    kw = dict(n_layers=spec["n_layers"], n_heads=spec["llm_heads"], eps=spec["eps"])
    if fix["kind"] == "llama":
        kw["theta"] = spec["rope_theta"]
    got = fwd(x, sd, **kw)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("kind", ["gpt2", "llama"])
def test_oracle_train_mode_backbone_dropout_matches_live_hf(kind, monkeypatch):
    """The frozen backbone's own dropouts are live in the reference's train mode (tasks/forecasting.py:18 flips the HF
    module): pin WHERE the oracle applies them — embd, attention probabilities, both residual branches — against
    HuggingFace itself run in train mode, by routing every torch dropout call of the HF forward through recorded masks
    and replaying the same masks in the oracle."""
    import transformers
    import torch.nn.functional as F
    torch.manual_seed(0)
    D, H, Lyr, L, B, p = 64, 2, 2, 10, 3, 0.25
    if kind == "gpt2":
        cfg = transformers.GPT2Config(n_embd=D, n_layer=Lyr, n_head=H, vocab_size=8, n_positions=32, attn_pdrop=p,
                                      embd_pdrop=p, resid_pdrop=p)
        hf = transformers.GPT2Model(cfg)
    else:
        cfg = transformers.LlamaConfig(hidden_size=D, intermediate_size=96, num_hidden_layers=Lyr, num_attention_heads=H,
                                       num_key_value_heads=H, vocab_size=8, rms_norm_eps=1e-5, attention_dropout=p)
        hf = transformers.LlamaModel(cfg)
    hf.config._attn_implementation = "eager"
    hf.train()
    x = torch.randn(B, L, D)
    calls = []
    g = torch.Generator().manual_seed(1)

    def recorded_dropout(inp, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return inp
        mask = (torch.rand(inp.shape, generator=g) >= p).to(inp.dtype) / (1 - p)
        calls.append(mask)
        return inp * mask

    monkeypatch.setattr(F, "dropout", recorded_dropout)
    monkeypatch.setattr(torch.nn.functional, "dropout", recorded_dropout)
    with torch.no_grad():
        ref = hf(inputs_embeds=x).last_hidden_state
    monkeypatch.undo()
    sd = {k: v for k, v in hf.state_dict().items()}
    if kind == "gpt2":
        assert len(calls) == 1 + 3 * Lyr                         # embd, then per block: attention probs, attn resid, mlp resid
        masks = {"embd": calls[0], "attn": calls[1::3], "resid_attn": calls[2::3], "resid_mlp": calls[3::3]}
        assert masks["attn"][0].shape == (B, H, L, L) and masks["resid_mlp"][1].shape == (B, L, D)
        got = O.gpt2_forward(x, sd, n_layers=Lyr, n_heads=H, dropout=masks)
    else:
        assert len(calls) == Lyr and calls[0].shape == (B, H, L, L)
        got = O.llama_forward(x, sd, n_layers=Lyr, n_heads=H, eps=1e-5, dropout={"attn": calls})
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("T,P,S", [(96, 16, 8), (100, 16, 8), (336, 16, 8), (512, 16, 8), (1024, 16, 8), (17, 16, 8), (64, 8, 4)])
def test_patch_index_is_unfold_of_replication_padded(T, P, S):
    """Bit-exact index contract (models/layers/embed.py:155-163,188-189; n_patches models/medtsllm.py:52)."""
    x = torch.arange(T, dtype=torch.float32)[None, None]
    padded = torch.cat([x, x[:, :, -1:].repeat(1, 1, S)], dim=-1)
    ref = padded.unfold(-1, P, S)[0, 0].long()
    idx = O.patch_index(T, P, S)
    assert torch.equal(idx, ref)
    assert idx.shape[0] == O.n_patches(T, P, S) == int((T - P) / S + 2)


def test_token_conv_matches_circular_conv1d():
    torch.manual_seed(0)
    p = torch.randn(6, 13, 16)
    conv = torch.nn.Conv1d(16, 32, 3, padding=1, padding_mode="circular", bias=False)
    ref = conv(p.permute(0, 2, 1)).transpose(1, 2)
    torch.testing.assert_close(O.token_conv(p, conv.weight.detach()), ref.detach(), rtol=1e-5, atol=1e-5)


def test_left_padding_uses_pad_embedding():
    emb = torch.randn(10, 4)
    out = O.assemble_prompt([[1, 2, 3], [4]], emb, pad_id=9)
    assert torch.equal(out[0], emb[[1, 2, 3]])
    assert torch.equal(out[1], emb[[9, 9, 4]])
    assert O.assemble_prompt([[], []], emb, 0).shape == (2, 0, 4)


def test_gelu_new_and_rope_tables():
    x = torch.linspace(-4, 4, 101)
    from transformers.activations import NewGELUActivation
    torch.testing.assert_close(O.gelu_new(x), NewGELUActivation()(x))
    cos, sin = O.rope_tables(7, 8)
    assert cos.shape == (7, 4) and torch.allclose(cos[0], torch.ones(4)) and torch.allclose(sin[0], torch.zeros(4))
    assert math.isclose(cos[1, 0].item(), math.cos(1.0), rel_tol=1e-6)


# ------------------------------------------------------------------------------------------------ GPT4TS
from _fixtures import GPT4TS_CASES, gpt4ts_spec, load_gpt4ts_backbone  # noqa: E402
from oracle import gpt4ts_oracle as G  # noqa: E402


@pytest.mark.parametrize("name", GPT4TS_CASES)
def test_gpt4ts_oracle_matches_reference_golden(name):
    """oracle/gpt4ts_oracle.py against tensors captured from the unmodified models/gpt4ts.py (BASELINE configs[0] =
    gpt4ts_forecast_etth1: seq_len = pred_len = 96, 7 variables, batch 8)."""
    fix = load_case(name)
    bb = load_gpt4ts_backbone()
    sd = {k: v.float() for k, v in bb["state"].items()}
    spec = gpt4ts_spec(fix, bb)
    out, st = G.gpt4ts_forward(fix["inputs"]["x_enc"], fix["params"], sd, spec, return_stages=True)
    g = fix["stages"]
    D = sd["wpe.weight"].shape[1]
    emb = torch.nn.functional.pad(st["embedding"], (0, D - st["embedding"].shape[-1]))
    torch.testing.assert_close(emb, g["gpt2_input"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(st["gpt2"], g["gpt2"], rtol=1e-4, atol=5e-5)
    assert out.shape == g["output"].shape
    torch.testing.assert_close(out, g["output"], rtol=1e-4, atol=5e-5)
    assert _rel_l2(out, g["output"]) < 2e-5
    out_t = G.gpt4ts_forward(fix["inputs"]["x_enc"], fix["params"], sd, spec, training=True)
    torch.testing.assert_close(out_t, g["output_train"], rtol=1e-4, atol=5e-5)


def test_gpt4ts_rejects_reconstruction_like_the_reference():
    with pytest.raises(ValueError):
        G.gpt4ts_forward(torch.zeros(1, 8, 2), {}, {}, dict(task="reconstruction"))


@pytest.mark.parametrize("name", ["llama_seg_concat", "gpt2_anomaly_concat", "llama_semseg_univariate"])
def test_reference_prompt_rows_are_identical_across_samples(name):
    """The premise of in-batch prompt sharing, checked on tensors captured from the UNMODIFIED reference:
    with a static dataset / task prompt the backbone is causal and unmasked, so at every layer the hidden states of the
    prompt positions are the same for every sample of the batch (the reference computes them once per sample anyway)."""
    fix = load_case(name)
    Lp = len(fix["prompt_ids"][0])
    assert all(p == fix["prompt_ids"][0] for p in fix["prompt_ids"])
    hs = fix["stages"]["llm.hidden_states"]
    assert len(hs) >= 2
    for h in hs:
        assert h.shape[0] >= 2
        for b in range(1, h.shape[0]):
            torch.testing.assert_close(h[b, :Lp], h[0, :Lp], rtol=1e-5, atol=1e-6)
        assert not torch.allclose(h[1, Lp:], h[0, Lp:])          # the patch rows do differ
