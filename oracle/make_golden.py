"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference on CPU.

Run in the build container only:   python -m oracle.make_golden
Each fixture is one self-contained file: the HF backbone config + bf16-representable random weights,
the synthetic tokenizer, the experiment config, the adapter state_dict (the reference's own init under
a fixed seed), the inputs, and the stage tensors captured by forward hooks on the reference modules.
Nothing at test time needs /root/reference.
"""
from __future__ import annotations

import zlib
import shutil
import sys
import tempfile
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import ref_harness as H  # noqa: E402

GOLDEN = REPO / "tests" / "golden"

CASES = {
    # BIDMC-like (BASELINE config 2): segmentation, 3 variables, concat, Llama
    "llama_seg_concat": dict(
        kind="llama", llm=dict(hidden_size=128, heads=2, layers=2, intermediate_size=256, vocab_size=384),
        task="segmentation", T=96, pred=96, C=3, B=4, num_tokens=256, d_ff=64, covariate_mode="concat",
        description="The BIDMC dataset contains PPG , ECG and respiration signals from ICU patients .",
        prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    # PSM-like (BASELINE config 4): anomaly detection = reconstruction + RevIN de-norm, GPT-2
    "gpt2_anomaly_concat": dict(
        kind="gpt2", llm=dict(hidden_size=128, heads=2, layers=2, vocab_size=384),
        task="anomaly_detection", T=100, pred=100, C=5, B=3, num_tokens=256, d_ff=64, covariate_mode="concat",
        description="PSM is a server machine dataset collected from application server nodes at eBay .",
        prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    # LUDB-like (BASELINE config 3): semantic segmentation, univariate, 4 classes, d_ff 128, head_dim 128
    "llama_semseg_univariate": dict(
        kind="llama", llm=dict(hidden_size=256, heads=2, layers=1, intermediate_size=256, vocab_size=384),
        task="semantic_segmentation", T=64, pred=64, C=1, n_classes=4, B=2, num_tokens=256, d_ff=128,
        covariate_mode="univariate",
        description="LUDB is an ECG signal database with marked boundaries of P , T waves and QRS complexes .",
        prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    # Ventilator-like (BASELINE config 5 without LoRA): forecasting with per-sample clip descriptions of
    # different lengths (exercises LEFT padding) and the input-statistics prompt
    "llama_forecast_clip_stats": dict(
        kind="llama", llm=dict(hidden_size=128, heads=2, layers=2, intermediate_size=320, vocab_size=384),
        task="forecasting", T=64, pred=24, C=2, B=3, num_tokens=256, d_ff=64, covariate_mode="concat",
        description="Ventilator pressure and flow waveforms recorded from ICU patients .",
        prompting=dict(dataset=True, task=True, clip=True, input_stats=True),
        descriptions=["Patient is sedated .", "Patient is awake and breathing with pressure support ventilation .",
                      "No notes ."]),
    # the two parameter-free down-sample modes (models/medtsllm.py:354-363), small on purpose
    "llama_forecast_truncate": dict(
        kind="llama", llm=dict(hidden_size=128, heads=2, layers=1, intermediate_size=256, vocab_size=256),
        task="forecasting", T=48, pred=16, C=2, B=2, num_tokens=64, d_ff=64, covariate_mode="concat",
        downsample="truncate", description="Synthetic two channel series .",
        prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    "gpt2_reconstruction_average": dict(
        kind="gpt2", llm=dict(hidden_size=128, heads=2, layers=1, vocab_size=256),
        task="reconstruction", T=40, pred=40, C=1, B=3, num_tokens=64, d_ff=32, covariate_mode="univariate",
        downsample="average", description="Synthetic single channel series .",
        prompting=dict(dataset=True, task=False, clip=False, input_stats=False)),
    # the remaining covariate modes (models/medtsllm.py:71-87, 284-295, 343-344, 369-377), small on purpose
    "llama_forecast_independent": dict(
        kind="llama", llm=dict(hidden_size=128, heads=2, layers=1, intermediate_size=256, vocab_size=256),
        task="forecasting", T=48, pred=16, C=3, B=2, num_tokens=64, d_ff=64, covariate_mode="independent",
        description="Synthetic three channel series .", prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    "gpt2_forecast_merge_end": dict(
        kind="gpt2", llm=dict(hidden_size=128, heads=2, layers=1, vocab_size=256),
        task="forecasting", T=48, pred=16, C=3, B=2, num_tokens=64, d_ff=64, covariate_mode="merge-end",
        description="Synthetic three channel series .", prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    "llama_anomaly_add": dict(
        kind="llama", llm=dict(hidden_size=128, heads=2, layers=1, intermediate_size=256, vocab_size=256),
        task="anomaly_detection", T=40, pred=40, C=3, B=2, num_tokens=64, d_ff=64, covariate_mode="add",
        description="Synthetic three channel series .", prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    "gpt2_anomaly_weighted_average": dict(
        kind="gpt2", llm=dict(hidden_size=128, heads=2, layers=1, vocab_size=256),
        task="anomaly_detection", T=40, pred=40, C=3, B=2, num_tokens=64, d_ff=64, covariate_mode="weighted-average",
        description="Synthetic three channel series .", prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
    # prompting.examples (models/medtsllm.py:402-405, :313-319; datasets/ecg.py:140-166): every sample carries
    # ("Example segment:", tensor [1, T_ex, C]) with its own T_ex -> ragged prompts, time-series rows inside the prompt
    "llama_seg_examples": dict(
        kind="llama", llm=dict(hidden_size=128, heads=2, layers=2, intermediate_size=256, vocab_size=384),
        task="segmentation", T=96, pred=96, C=2, B=3, num_tokens=128, d_ff=64, covariate_mode="concat",
        description="ECG recordings with annotated beats .", example_lens=[40, 56, 24],
        prompting=dict(dataset=True, task=True, clip=False, input_stats=False, examples=True)),
    "llama_forecast_interleave": dict(
        kind="llama", llm=dict(hidden_size=128, heads=2, layers=1, intermediate_size=256, vocab_size=256),
        task="forecasting", T=48, pred=16, C=3, B=2, num_tokens=64, d_ff=64, covariate_mode="interleave",
        description="Synthetic three channel series .", prompting=dict(dataset=True, task=True, clip=False, input_stats=False)),
}


def _bf16_representable_(model):
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(p.to(torch.bfloat16).float())


def make_case(name: str, c: dict, out_path: Path):
    import transformers
    tmp = Path(tempfile.mkdtemp(prefix="mts_golden_"))
    try:
        ds = H.SyntheticDataset(c["C"], n_classes=c.get("n_classes", 0), description=c["description"])
        # tokenizer vocabulary: every word that can appear in this case's prompts
        texts = [f"Dataset: {c['description']}", "Task: Forecast the next steps given the previous steps of data .",
                 "Reconstruct the past steps of data as accurately as possible using the following information .",
                 "Classify Identify the change points in to segment sequence . Time series:",
                 "Input statistics ( feature 0 ): min value = , max median the trend of input is upward downward top 5 lags are"]
        texts += c.get("descriptions", []) + ["Example segment:"]
        texts += [str(c["T"]), str(c["pred"])]
        llm_dir = H.build_llm_dir(c["kind"], tmp / "llm", texts, seed=zlib.crc32(name.encode()) % 1000, **c["llm"])
        # round the backbone to bf16-representable values so that weight rounding is not part of the
        # parity error budget (activations still are)
        cls = transformers.LlamaModel if c["kind"] == "llama" else transformers.GPT2Model
        hf = cls.from_pretrained(str(llm_dir), torch_dtype=torch.float32)
        _bf16_representable_(hf)
        hf.save_pretrained(str(llm_dir))

        prompting = dict(H.make_config(task="x", history_len=1, pred_len=1, llm_path="")["models"]["medtsllm"]["prompting"])
        prompting.update(c["prompting"])
        cfg = H.make_config(task=c["task"], history_len=c["T"], pred_len=c["pred"], llm_path=llm_dir,
                            d_ff=c["d_ff"], num_tokens=c["num_tokens"], covariate_mode=c["covariate_mode"],
                            downsample=c.get("downsample", "linear"), prompting=prompting)
        model = H.build_reference_model(cfg, ds, seed=1)

        g = torch.Generator().manual_seed(1234)
        z = torch.randn(c["B"], c["T"], c["C"], generator=g)
        scale = torch.rand(c["C"], generator=g) * 4.5 + 0.5
        offset = torch.rand(c["C"], generator=g) * 20 - 10
        inputs = {"x_enc": z * scale + offset}
        if "descriptions" in c:
            inputs["descriptions"] = list(c["descriptions"])
        if "example_lens" in c:      # what datasets/ecg.py's collate_fn hands over: (text, tensor [1, T_ex, C]) per sample
            inputs["examples"] = [("Example segment:", (torch.randn(1, n, c["C"], generator=g) * scale + offset))
                                  for n in c["example_lens"]]
        out, stages = H.run_reference_with_stages(model, inputs)
        prompts = model.build_prompt(inputs)
        # flat per-sample lists: token ids, with time-series example parts kept as tensors in their place
        prompt_ids = [[t for part in parts for t in ([part] if isinstance(part, torch.Tensor) else
                                                     model.tokenizer(part, padding=False, truncation=False).input_ids)]
                      for parts in prompts]
        # train-mode forward (dropout 0): no eval-only activation — the tensor the loss sees
        out_train, _ = H.run_reference_with_stages(model, inputs, train_mode=True)

        cfg["models"]["medtsllm"]["llm"]["llm"] = "<llm_dir>"   # filled in at test time
        fixture = {
            "name": name,
            "kind": c["kind"],
            "hf_config": hf.config.to_dict(),
            "backbone_state": {k: v.to(torch.bfloat16) for k, v in hf.state_dict().items()},
            "tokenizer_json": (llm_dir / "tokenizer.json").read_text(),
            "tokenizer_bos": c["kind"] == "llama",
            "config": cfg,
            "dataset": {"n_features": c["C"], "n_classes": c.get("n_classes", 0), "description": c["description"]},
            "inputs": inputs,
            "adapters": {k: v.clone() for k, v in model.state_dict().items()},
            "prompts": prompts,
            "prompt_ids": prompt_ids,
            "pad_id": model.tokenizer.pad_token_id,
            "stages": {
                "patch_embedding": stages["patch_embedding"],
                "source_embeddings": stages["mapping_layer"].permute(1, 0).contiguous(),
                "reprogramming_layer": stages["reprogramming_layer"],
                "llm_input": stages["llm_input"],
                "llm": stages["llm"],
                "llm.hidden_states": stages["llm.hidden_states"],
                "downsample": stages.get("embedding_downsample_layer"),
                "output_projection": stages["output_projection"],
                "revin_mean": stages["revin_mean"],
                "revin_stdev": stages["revin_stdev"],
                "output": stages["output"],
                "output_train": out_train.detach().clone(),
            },
            "generator": "oracle/make_golden.py (reference e873455, transformers " + transformers.__version__ + ")",
        }
        torch.save(fixture, out_path)
        print(f"{name}: wrote {out_path} ({out_path.stat().st_size / 1e6:.2f} MB), output {tuple(out.shape)}, "
              f"Lp={[len(p) for p in prompt_ids]}, llm_input {tuple(stages['llm_input'].shape)}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        make_case(name, c, GOLDEN / f"{name}.pt")


if __name__ == "__main__":
    main()
