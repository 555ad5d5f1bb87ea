"""TEST INFRASTRUCTURE — generates tests/golden/gpt4ts_*.pt by running the UNMODIFIED reference class
`models.gpt4ts.GPT4TS` on CPU (BASELINE.json configs[0]: ETTh1 forecasting on a frozen GPT-2).

Run in the build container only:   python -m oracle.make_golden_gpt4ts
The reference hard-codes `GPT2Model.from_pretrained("gpt2")` and a 768-wide residual stream
(models/gpt4ts.py:44, :138); no checkpoint exists offline, so a seeded random-init 768-wide GPT-2 is saved to
./gpt2 of a scratch working directory and the reference loads that.  The backbone weights (bf16-representable)
are stored once (gpt4ts_backbone.pt) and shared by the cases.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import tempfile
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import ref_harness as H  # noqa: E402

GOLDEN = REPO / "tests" / "golden"
GPT_LAYERS = 2

CASES = {
    # BASELINE configs[0]: ETTh1 forecasting, seq_len = pred_len = 96, 7 variables, batch 8
    "gpt4ts_forecast_etth1": dict(task="forecasting", T=96, pred=96, C=7, B=8),
    "gpt4ts_anomaly": dict(task="anomaly_detection", T=100, pred=100, C=5, B=3),
    "gpt4ts_semseg": dict(task="semantic_segmentation", T=64, pred=64, C=2, B=2, n_classes=4),
    "gpt4ts_segmentation": dict(task="segmentation", T=48, pred=48, C=1, B=2),
}


def make_config(c):
    return {
        "task": c["task"], "model": "gpt4ts", "history_len": c["T"], "pred_len": c["pred"],
        "training": {"dropout": 0.0, "batch_size": c["B"], "learning_rate": 1e-4},
        "setup": {"dtype": "float32", "seed": 0},
        "tasks": {"segmentation": {"mode": "boundary-prediction"}},
        "models": {"gpt4ts": {"d_ff": 768, "d_model": 768, "gpt_layers": GPT_LAYERS, "train_mlp": False,
                              "patching": {"patch_len": 1, "stride": 1}}},
    }


def main():
    import transformers
    os.environ["HF_HUB_OFFLINE"] = "1"
    ref = H.import_reference()
    gpt4ts = importlib.import_module("models.gpt4ts")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    tmp = Path(tempfile.mkdtemp(prefix="mts_gpt4ts_"))
    cwd = os.getcwd()
    try:
        torch.manual_seed(11)
        # n_inner = 768 keeps the stored backbone small; one more block than gpt_layers exercises the slicing (:45)
        cfg = transformers.GPT2Config(vocab_size=64, n_embd=768, n_layer=GPT_LAYERS + 1, n_head=12, n_positions=256,
                                      n_inner=768, attn_pdrop=0.0, embd_pdrop=0.0, resid_pdrop=0.0,
                                      bos_token_id=2, eos_token_id=2)
        hf = transformers.GPT2Model(cfg)
        with torch.no_grad():
            for p in hf.parameters():
                p.copy_(p.to(torch.bfloat16).float())
            # LayerNorm / bias / position parameters are trainable in GPT4TS (:47-53): give them non-trivial values
            g = torch.Generator().manual_seed(12)
            for n, p in hf.named_parameters():
                if "ln" in n or n.endswith(".bias"):
                    p.copy_((p + torch.randn(p.shape, generator=g) * 0.05).to(torch.bfloat16).float())
        (tmp / "gpt2").mkdir()
        hf.save_pretrained(str(tmp / "gpt2"))
        os.chdir(tmp)
        keep = lambda k: not k.startswith("h.") or int(k.split(".")[1]) < GPT_LAYERS     # noqa: E731
        torch.save({"hf_config": cfg.to_dict(),
                    "state": {k: v.to(torch.bfloat16) for k, v in hf.state_dict().items() if keep(k) and not k.startswith("wte")}},
                   GOLDEN / "gpt4ts_backbone.pt")
        only = sys.argv[1:]
        for name, c in CASES.items():
            if only and name not in only:
                continue
            ds = H.SyntheticDataset(c["C"], n_classes=c.get("n_classes", 0), description="")
            config = make_config(c)
            torch.manual_seed(1)
            with H._quiet():
                model = gpt4ts.GPT4TS(ref.dict_to_object(config), ds)
            g = torch.Generator().manual_seed(1234)
            z = torch.randn(c["B"], c["T"], c["C"], generator=g)
            x = z * (torch.rand(c["C"], generator=g) * 4.5 + 0.5) + (torch.rand(c["C"], generator=g) * 20 - 10)
            stages = {}
            hook = model.gpt2.register_forward_hook(
                lambda _m, _i, out: stages.__setitem__("gpt2", out.last_hidden_state.detach().clone()))
            pre = model.gpt2.register_forward_pre_hook(
                lambda _m, args, kwargs: stages.__setitem__("gpt2_input", kwargs["inputs_embeds"].detach().clone()),
                with_kwargs=True)
            model.eval()
            with torch.no_grad():
                out = model({"x_enc": x.clone()})           # (the reference normalises x_enc in place)
            model.train()
            with torch.no_grad():
                out_train = model({"x_enc": x.clone()})
            hook.remove(); pre.remove()
            own = {k: v.clone() for k, v in model.state_dict().items()
                   if not k.startswith("gpt2.") and not k.endswith("position_embedding.pe")}
            fixture = {"name": name, "config": config,
                       "dataset": {"n_features": c["C"], "n_classes": c.get("n_classes", 0), "description": ""},
                       "inputs": {"x_enc": x}, "params": own,
                       "stages": {"gpt2_input": stages["gpt2_input"], "gpt2": stages["gpt2"],
                                  "output": out.detach().clone(), "output_train": out_train.detach().clone()},
                       "generator": "oracle/make_golden_gpt4ts.py (unmodified models/gpt4ts.py, transformers "
                                    + transformers.__version__ + ")"}
            path = GOLDEN / f"{name}.pt"
            torch.save(fixture, path)
            print(f"{name}: wrote {path} ({path.stat().st_size / 1e6:.2f} MB), output {tuple(out.shape)}, "
                  f"params {sorted(own)}")
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
