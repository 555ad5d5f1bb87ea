"""The reference Trainer's call sequence on the kernel stack, on the GPU.

/root/reference does not exist on the GPU box, so the loop is RESTATED here line by line (each block cites the lines it
follows); the same sequence is driven through the reference's real `tasks.ForecastTask` in tests/test_reference_trainer.py
(CPU, build container), where it stops at the first forward for want of a GPU.  Covered:

  * tasks/base.py:27-55      constructor order: loaders -> model.to(device, dtype) -> Adam over requires_grad params
  * tasks/base.py:163-198    DataLoaders: shuffle, pin_memory=True, worker processes
  * tasks/base.py:200-211    prepare_batch (+ the plugin's HostMirrorBatch in evaluation)
  * tasks/forecasting.py:15-36   train(): model.train(), bf16 autocast, loss, backward, step, zero_grad, loss.item(),
                                 val() after every epoch (weights changed -> cached bf16 copies / captured graphs refresh)
  * tasks/forecasting.py:52-95   predict(): eval, per-sample scatter through dataset.inverse_index, ragged last batch
  * loggers/base_logger.py:29-43 save_state (+ LoRA pairs) and tasks/base.py:283-306 from_run_id (+ the plugin's LoRA reload)
  * tasks/base.py:277-281    SIGUSR1 -> state_dict() between two bytecodes of a training step
"""
import os
import signal
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset, default_collate

from _fixtures import Cfg, load_case, materialize_llm_dir

pytestmark = pytest.mark.gpu

HIST, PRED, STEP, BATCH = 64, 16, 16, 4


class SynthForecast(Dataset):
    """datasets/base.py:107-137 (ForecastDataset): windows of a [n_points, C] series."""
    univariate = False
    clip_dataset = False

    def __init__(self, split):
        g = torch.Generator().manual_seed({"train": 0, "val": 1, "test": 2}[split])
        n = 306 if split == "train" else 178            # 14 train windows -> ragged last batch (4 + 4 + 4 + 2)
        t = torch.arange(n, dtype=torch.float32)
        self.data = torch.stack([torch.randn(n, generator=g).cumsum(0) * 0.3, torch.sin(t / 7.0) + 0.1 * torch.randn(n, generator=g)], 1)
        self.step_size = PRED if split == "test" else STEP
        self.n_features, self.n_classes, self.description, self.task_description = 2, 0, "Two synthetic channels .", None

    n_points = property(lambda self: self.data.shape[0])
    real_features = property(lambda self: self.n_features)

    def __len__(self):
        return (self.n_points - HIST - PRED + 1) // self.step_size

    def inverse_index(self, idx):
        idx = idx * self.step_size
        return (idx, idx + HIST), (idx + HIST, idx + HIST + PRED)

    def __getitem__(self, idx):
        xr, yr = self.inverse_index(idx)
        return {"x_enc": self.data[slice(*xr)], "y": self.data[slice(*yr)]}


class MirrorForecastTask:
    """Restatement of tasks/base.py BaseTask + tasks/forecasting.py ForecastTask (the parts the hot path sees)."""

    def __init__(self, run_id, config, logdir, newrun=True):
        from medtsllm_b200.model import MedTsLLM
        self.run_id, self.config, self.logdir = run_id, config, Path(logdir)
        self.device, self.dtype, self.mixed = torch.device("cuda"), torch.float32, config.setup.dtype == "mixed"
        torch.manual_seed(config.setup.seed)                                            # base.py:36
        self.train_dataset, self.val_dataset = SynthForecast("train"), SynthForecast("val")
        kw = dict(batch_size=config.training.batch_size, collate_fn=default_collate, num_workers=2, pin_memory=True)
        self.train_dataloader = DataLoader(self.train_dataset, shuffle=True, **kw)      # base.py:175-182
        self.val_dataloader = DataLoader(self.val_dataset, shuffle=False, **kw)         # base.py:183-190
        self.model = MedTsLLM(config, self.train_dataset).to(self.device, self.dtype)   # base.py:41, :81-85
        params = [p for p in self.model.parameters() if p.requires_grad]                # base.py:93
        self.optimizer = torch.optim.Adam(params, lr=config.training.learning_rate)     # base.py:97
        self.loss_fn = torch.nn.MSELoss().to(self.device)
        self.epoch, self.step, self.losses = 1, 0, []

    def prepare_batch(self, batch):                                                      # base.py:200-211
        if isinstance(batch, dict):
            return {k: self.prepare_batch(v) for k, v in batch.items()}
        if isinstance(batch, (list, tuple)):
            return [self.prepare_batch(x) for x in batch]
        if isinstance(batch, torch.Tensor):
            batch = batch.to(self.device)
            return batch.to(self.dtype) if batch.dtype.is_floating_point else batch
        return batch

    def train(self, epochs, on_step=None):                                              # forecasting.py:15-36
        self.val_scores = []
        for _ in range(epochs):
            self.model.train()
            for inputs in self.train_dataloader:
                inputs = self.prepare_batch(inputs)
                with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=self.mixed):
                    pred = self.model(inputs)
                    loss = self.loss_fn(pred, inputs["y"])
                loss.backward()
                if on_step is not None:
                    on_step()
                self.optimizer.step()
                self.optimizer.zero_grad()
                self.losses.append(loss.item())
                self.step += self.config.training.batch_size
            preds, targets = self.predict(self.val_dataloader)                           # val(): forecasting.py:38-43
            self.val_scores.append(F.mse_loss(preds, targets).item())
            self.save_state("latest")                                                    # base.py:229 log_epoch
            self.epoch += 1
        self.model.eval()

    def predict(self, dataloader):                                                       # forecasting.py:52-95
        self.model.eval()
        dataset = dataloader.dataset
        pred_len, ctx_len, step_size = self.config.pred_len, self.config.history_len, dataset.step_size
        n_points = pred_len + ctx_len + ((len(dataset) - 1) * step_size)
        n_features, bs = dataset.real_features, dataloader.batch_size
        preds = torch.full((n_points, n_features), float("nan"))
        targets = torch.full((n_points, n_features), float("nan"))
        with torch.no_grad():
            for idx, inputs in enumerate(dataloader):
                inputs = self.prepare_batch(inputs)
                pred = self.model(inputs)
                for j in range(pred.size(0)):
                    inds = dataset.inverse_index((idx * bs) + j)
                    time_inds = slice(*inds[1])
                    preds[time_inds, slice(None)] = pred[j].squeeze().cpu().detach()
                    targets[time_inds, slice(None)] = inputs["y"][j].squeeze().cpu().detach()
        preds, targets = preds[ctx_len:], targets[ctx_len:]
        assert not preds.isnan().any() and not targets.isnan().any()
        return preds, targets

    def save_state(self, name):                                                          # base_logger.py:29-43
        d = self.logdir / self.run_id / "checkpoints"
        d.mkdir(parents=True, exist_ok=True)
        torch.save({"run_id": self.run_id, "epoch": self.epoch, "step": self.step, "model": self.model.state_dict()},
                   d / f"{name}.pt")
        if hasattr(self.model, "lora_enabled") and self.model.lora_enabled:
            self.model.llm.save_pretrained(d / f"{name}-lora.safetensors")

    @classmethod
    def from_run_id(cls, run_id, cfg=None, ckpt="latest", basepath=None):               # base.py:283-306
        trainer = cls(run_id, cfg, basepath, newrun=False)
        state = torch.load(Path(basepath) / run_id / f"checkpoints/{ckpt}.pt")
        _, unexpected = trainer.model.load_state_dict(state["model"], strict=False)
        assert not unexpected
        trainer.epoch, trainer.step = state["epoch"], state["step"]
        return trainer


def _config(llm_dir, lora=False, dropout=0.0):
    return Cfg({
        "task": "forecasting", "model": "medtsllm", "history_len": HIST, "pred_len": PRED,
        "training": {"epochs": 2, "batch_size": BATCH, "optimizer": "adam", "learning_rate": 1e-3, "dropout": dropout, "loss": "mse"},
        "tasks": {"segmentation": {"mode": "boundary-prediction"}},
        "models": {"medtsllm": {
            "d_model": 32, "d_ff": 64, "n_heads": 8, "num_tokens": 64, "covariate_mode": "concat",
            "embedding_downsample_mode": "linear", "patching": {"patch_len": 16, "stride": 8},
            "prompting": {"dataset": True, "task": True, "clip": False, "input_stats": False, "examples": False,
                          "input_stats_dim": 0, "input_stats_select": "all"},
            "llm": {"enabled": True, "llm": str(llm_dir), "llm_layers": -1, "load_in_4bit": False, "load_in_8bit": False},
            **({"lora": {"enabled": True, "layers": "auto", "rank": 4, "alpha": 8, "init": True, "dropout": 0.0,
                         "rslora": True}} if lora else {}),
        }},
        "setup": {"seed": 0, "device": "auto", "dtype": "mixed", "num_workers": 2, "logger": "print"},
    })


@pytest.fixture()
def patched_task():
    """MirrorForecastTask with the plugin's Trainer wrappers applied (the same function patches tasks.base.BaseTask)."""
    from medtsllm_b200 import plugin

    class Task(MirrorForecastTask):
        pass
    return plugin.patch_trainer(Task)


@pytest.mark.parametrize("lora", [False, True])
def test_trainer_sequence_end_to_end(lora, tmp_path, cuda, patched_task):
    llm_dir = materialize_llm_dir(load_case("llama_forecast_truncate"), tmp_path / "llm")
    cfg = _config(llm_dir, lora=lora)
    trainer = patched_task("run-gpu", cfg, tmp_path / "logs")
    model = trainer.model
    assert len(trainer.train_dataloader) == 4 and len(trainer.train_dataset) == 14       # ragged last batch of 2
    before = {k: v.clone() for k, v in model.state_dict().items()}
    hits = {"n": 0, "finite": True}

    def sigusr1(signum, frame):            # base.py:277-281: handle_termination -> logger.save_state -> state_dict()
        sd = model.state_dict()
        hits["n"] += 1
        hits["finite"] &= all(torch.isfinite(v).all().item() for v in sd.values())

    old = signal.signal(signal.SIGUSR1, sigusr1)
    try:
        trainer.train(2, on_step=lambda: os.kill(os.getpid(), signal.SIGUSR1))           # mid-step, between backward and step
    finally:
        signal.signal(signal.SIGUSR1, old)
    assert hits["n"] == 8 and hits["finite"]
    assert len(trainer.losses) == 8 and all(l == l and l < 1e4 for l in trainer.losses)
    assert sum(trainer.losses[4:]) < sum(trainer.losses[:4])                             # it learns (same 14 windows, lr 1e-3)
    after = model.state_dict()
    assert all(not torch.equal(before[k], after[k]) for k in before if not k.endswith("key_projection.bias"))
    assert trainer.val_scores[0] != trainer.val_scores[1]                                # val() saw the updated weights

    # evaluation: plain device batches vs the plugin's host-mirror batches — same numbers, predictions already on the host
    os.environ["MTS_HOST_MIRROR"] = "0"
    try:
        p_plain, t_plain = trainer.predict(trainer.val_dataloader)
    finally:
        os.environ["MTS_HOST_MIRROR"] = "1"
    p_mirror, t_mirror = trainer.predict(trainer.val_dataloader)
    assert torch.equal(p_plain, p_mirror) and torch.equal(t_plain, t_mirror)
    batch = next(iter(trainer.val_dataloader))
    model.eval()
    hm = trainer.prepare_batch(batch)
    assert type(hm).__name__ == "HostMirrorBatch" and hm["x_enc"].device.type == "cpu" and hm.device_batch["x_enc"].is_cuda
    with torch.no_grad():
        assert model(hm).device.type == "cpu"
    model.train()
    assert type(trainer.prepare_batch(batch)) is dict                                    # training batches are untouched
    # repeated evaluation replays the captured graph: still the same numbers
    for _ in range(3):
        assert torch.equal(trainer.predict(trainer.val_dataloader)[0], p_mirror)

    # checkpoint -> fresh process-like reload (incl. LoRA pairs through the plugin's from_run_id wrapper)
    reloaded = patched_task.from_run_id("run-gpu", cfg=cfg, ckpt="latest", basepath=tmp_path / "logs")
    assert reloaded.epoch == 2 and reloaded.step == 2 * 16
    a, b = dict(model.named_parameters()), dict(reloaded.model.named_parameters())
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(reloaded.predict(reloaded.val_dataloader)[0], p_mirror)


def test_weight_update_through_param_data_is_seen(tmp_path, cuda):
    """in-place updates through `p.data` do not bump `_version`; an optimizer's step() (global hook) or
    `invalidate_caches()` must still refresh the bf16 copies and captured graphs."""
    from medtsllm_b200.model import MedTsLLM
    llm_dir = materialize_llm_dir(load_case("llama_forecast_truncate"), tmp_path / "llm")
    ds = SynthForecast("val")
    model = MedTsLLM(_config(llm_dir), ds).to(cuda, torch.float32).eval()
    x = torch.stack([ds[i]["x_enc"] for i in range(4)]).to(cuda)
    with torch.no_grad():
        outs = [model({"x_enc": x}).clone() for _ in range(3)]                           # eager, capture, replay
        assert torch.equal(outs[0], outs[2])
        w = model.output_projection.linear.weight
        v0 = w._version
        w.data.mul_(1.5)                                                                 # weight surgery behind autograd's back
        assert w._version == v0
        model.invalidate_caches()
        o2 = model({"x_enc": x}).clone()
        assert not torch.equal(o2, outs[0])

    class DataSGD(torch.optim.Optimizer):                                                # an optimizer that updates p.data
        def __init__(self, params):
            super().__init__(params, {})

        def step(self):
            for g in self.param_groups:
                for p in g["params"]:
                    p.data.mul_(0.5)

    DataSGD([w]).step()
    with torch.no_grad():
        o3 = model({"x_enc": x}).clone()
    assert not torch.equal(o3, o2)
