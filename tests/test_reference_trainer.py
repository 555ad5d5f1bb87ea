"""Drop-in proof against the reference's REAL Trainer (build container only: needs /root/reference; skipped elsewhere).

The unmodified `tasks.get_trainer` / `ForecastTask.__init__` (tasks/base.py:27-55: datasets, DataLoaders, `build_model()`
through `models.model_lookup`, `.to(device, dtype)`, `build_optimizer`, scheduler, loss, logger, SIGUSR1 handler) runs with
`medtsllm_b200.plugin.register()` + `patch_trainer()` applied, on a synthetic dataset registered in the reference's own
`datasets.dataset_lookup`.  This container has no GPU, so the first forward of `trainer.train()` must fail with exactly
the no-CPU-fallback `MtsError` — and nothing before it.  Checkpoint round trip (`logger.save_state`,
`BaseTask.from_run_id`, LoRA pairs included) needs no forward and is checked here too.  The GPU half of the same call
sequence is tests/test_trainer_gpu.py (a line-by-line restatement of the loop, since the reference does not travel)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import ref_harness as H  # noqa: E402
from _fixtures import load_case, materialize_llm_dir  # noqa: E402

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="/root/reference only exists in the build container")


def _trainer_config(llm_dir, logdir, *, lora=False, task="forecasting"):
    return {
        "DEBUG": False, "task": task, "model": "medtsllm", "history_len": 64, "pred_len": 16,
        "data": {"dataset": "synthetic", "mode": "multivariate", "cols": "all", "normalize": True, "step": 16},
        "training": {"epochs": 1, "batch_size": 4, "optimizer": "adam", "learning_rate": 1e-4, "dropout": 0.0,
                     "loss": "mse", "eval_metric": "mse", "eval_metric_direction": "min"},
        "tasks": {"segmentation": {"mode": "boundary-prediction"}},
        "paths": {"logdir": str(logdir)},
        "models": {"medtsllm": {
            "d_model": 32, "d_ff": 64, "n_heads": 8, "num_tokens": 64, "covariate_mode": "concat",
            "embedding_downsample_mode": "linear", "patching": {"patch_len": 16, "stride": 8},
            "prompting": {"dataset": True, "task": True, "clip": False, "input_stats": False, "examples": False,
                          "input_stats_dim": 0, "input_stats_select": "all"},
            "llm": {"enabled": True, "llm": str(llm_dir), "llm_layers": -1, "load_in_4bit": False, "load_in_8bit": False},
            **({"lora": {"enabled": True, "layers": "auto", "rank": 4, "alpha": 8, "init": True, "dropout": 0.0,
                         "rslora": True}} if lora else {}),
        }},
        "setup": {"seed": 0, "device": "auto", "dtype": "mixed", "num_workers": 0, "logger": "print"},
    }


@pytest.fixture(scope="module")
def ref():
    ns = H.import_trainer()
    from medtsllm_b200 import plugin
    plugin.register(str(H.REFERENCE))
    plugin.patch_trainer(ns.tasks_base.BaseTask)

    class SyntheticForecast(ns.datasets_base.ForecastDataset):
        """Two synthetic channels: a random walk and a noisy sine."""
        supported_tasks = ["forecasting"]

        def get_data(self, split=None):
            split = split or self.split
            g = np.random.default_rng({"train": 0, "val": 1, "test": 2}[split])
            n = 400 if split == "train" else 180
            t = np.arange(n)
            return {"data": np.stack([np.cumsum(g.normal(size=n)), np.sin(t / 7.0) + 0.1 * g.normal(size=n)], axis=1)}

    ns.datasets.dataset_lookup["synthetic"] = {"forecasting": SyntheticForecast}
    return ns


def test_reference_trainer_builds_our_model_and_stops_only_at_the_gpu(ref, tmp_path, capsys):
    from medtsllm_b200._lib import MtsError
    from medtsllm_b200.model import MedTsLLM
    llm_dir = materialize_llm_dir(load_case("llama_forecast_truncate"), tmp_path / "llm")
    cfg = ref.dict_to_object(_trainer_config(llm_dir, tmp_path / "logs"))
    trainer = ref.tasks.get_trainer("run-cpu", cfg)                       # the reference's own constructor chain
    assert type(trainer).__name__ == "ForecastTask" and isinstance(trainer.model, MedTsLLM)
    assert trainer.device.type == "cpu" and trainer.dtype == torch.float32 and trainer.mixed
    # what build_optimizer saw (tasks/base.py:87-93): exactly the adapter tensors, all trainable fp32 masters
    opt_params = [p for g in trainer.optimizer.param_groups for p in g["params"]]
    named = dict(trainer.model.named_parameters())
    assert len(opt_params) == len(named) == 15 and all(p.requires_grad and p.dtype == torch.float32 for p in opt_params)
    assert len(trainer.train_dataloader) > 0 and (tmp_path / "logs" / "run-cpu" / "config.toml").exists()
    # first forward of the reference's own train loop: the ONLY failure is the missing GPU
    with pytest.raises(MtsError, match="CUDA device only"):
        trainer.train()
    # ... and it failed inside the model call of tasks/forecasting.py:23, after prepare_batch and autocast were entered
    assert trainer.step == 0


@pytest.mark.parametrize("lora", [False, True])
def test_checkpoint_round_trip_through_logger_and_from_run_id(ref, tmp_path, lora):
    """loggers/base_logger.py:29-43 save_state -> tasks/base.py:283-306 from_run_id, with the plugin's LoRA reload."""
    llm_dir = materialize_llm_dir(load_case("llama_forecast_truncate"), tmp_path / "llm")
    cfg = ref.dict_to_object(_trainer_config(llm_dir, tmp_path / "logs", lora=lora))
    trainer = ref.tasks.get_trainer("run-ckpt", cfg)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in trainer.model.parameters():                              # stand-in for training: perturb every tensor
            p.add_(torch.randn(p.shape, generator=g) * 0.01)
    trainer.epoch, trainer.step = 3, 123
    trainer.logger.save_state("latest")
    ckdir = tmp_path / "logs" / "run-ckpt" / "checkpoints"
    state = torch.load(ckdir / "latest.pt")
    assert set(state["model"]) == set(trainer.model.state_dict()) and not any(k.startswith("llm.") for k in state["model"])
    assert (ckdir / "latest-lora.safetensors").exists() == lora
    cls = type(trainer)
    reloaded = cls.from_run_id("run-ckpt", ckpt="latest", basepath=tmp_path / "logs")
    assert reloaded.epoch == 3 and reloaded.step == 123
    a, b = dict(trainer.model.named_parameters()), dict(reloaded.model.named_parameters())
    assert a.keys() == b.keys()
    for k in a:                                                           # adapters AND (plugin) the LoRA pairs
        assert torch.equal(a[k], b[k]), k


def test_sigusr1_handler_saves_state_between_bytecodes(ref, tmp_path):
    """tasks/base.py:277-281: the SIGUSR1 handler calls logger.save_state (model.state_dict()) wherever the main thread is."""
    import signal
    llm_dir = materialize_llm_dir(load_case("llama_forecast_truncate"), tmp_path / "llm")
    cfg = ref.dict_to_object(_trainer_config(llm_dir, tmp_path / "logs"))
    trainer = ref.tasks.get_trainer("run-sig", cfg)
    with pytest.raises(SystemExit):
        os.kill(os.getpid(), signal.SIGUSR1)
        for _ in range(1000):          # the handler runs between two bytecodes of this loop
            pass
    assert (tmp_path / "logs" / "run-sig" / "checkpoints" / "latest.pt").exists()
    signal.signal(signal.SIGUSR1, signal.SIG_DFL)
