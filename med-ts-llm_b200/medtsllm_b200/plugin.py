"""Drop-in registration into the reference (flixpar/med-ts-llm) — the plug-in seam of SURVEY.md §8b.

The reference builds its model with `model_lookup[config.model](config, train_dataset)`
(tasks/base.py:81-85, models/__init__.py:10-18).  `register()` points the keys "medtsllm" and
"timellm" at `medtsllm_b200.MedTsLLM`; `train.py`, `tasks/*`, `datasets/*` and `loggers/*` then run
unchanged:

    python -m medtsllm_b200.plugin /path/to/med-ts-llm/train.py configs/datasets/bidmc.toml
    torchrun --nproc-per-node 8 -m medtsllm_b200.plugin /path/to/med-ts-llm/train.py cfg.toml   # DP

`patch_trainer()` (called by the launcher) wraps three `BaseTask` methods without touching tasks/*:
  * `from_run_id` (tasks/base.py:283-306) also restores the LoRA pairs that `logger.save_state` wrote next to the
    checkpoint (loggers/base_logger.py:42-43) — the reference saves them but never loads them back;
  * `prepare_batch` (tasks/base.py:200-211), evaluation only: returns a `HostMirrorBatch` — the pinned host tensors the
    DataLoader produced, with their device copies attached.  The model computes on the device copies and returns its
    predictions on the host (ONE device->host copy per batch), so the per-sample `.cpu()` calls of the reference's
    `predict()` scatter loops (tasks/forecasting.py:72-78, anomaly_detection.py:106-113, semantic_segmentation.py:98-104,
    segmentation.py:92-97: up to 3 B device syncs per batch) all become no-ops.  Same values, same loop.
    `MTS_HOST_MIRROR=0` turns it off.
  * with WORLD_SIZE > 1, `__init__`: NCCL init, `train_dataloader` swapped for a DistributedSampler-backed one that
    reshuffles every epoch (the reference has `shuffle=True` and no sampler, tasks/base.py:175-182), real logging on
    rank 0 only.  Gradient all-reduce happens inside the model's backward (medtsllm_b200/dp.py).
"""
from __future__ import annotations

import os
import runpy
import sys
from pathlib import Path


def register(reference_root: str | None = None):
    """Mutates the reference's `models.model_lookup` in place.  Import order matters: this must run
    before `tasks` is imported by train.py, which is why the launcher below calls it first."""
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import models  # the reference's package (sys.path[0] is the reference root when train.py runs)
    from .model import MedTsLLM
    models.model_lookup["medtsllm"] = MedTsLLM
    models.model_lookup["timellm"] = MedTsLLM
    if os.environ.get("MTS_REGISTER_GPT4TS", "0") == "1":
        # opt-in: medtsllm_b200.GPT4TS keeps only its trainable tensors in state_dict() (not the frozen GPT-2
        # weights), so checkpoints written by the reference's own GPT4TS do not load into it
        from .gpt4ts import GPT4TS
        models.model_lookup["gpt4ts"] = GPT4TS
    return models.model_lookup


class HostMirrorBatch(dict):
    """The batch as the DataLoader produced it (host tensors, pinned when `pin_memory=True`) with the device copies
    `BaseTask.prepare_batch` made attached as `.device_batch`.  Indexing it yields HOST tensors."""

    device_batch: dict


def lora_checkpoint_path(basepath, run_id, ckpt="latest") -> Path:
    """Where loggers/base_logger.py:29-43 writes the LoRA pairs of checkpoint `ckpt`."""
    return Path(basepath) / run_id / "checkpoints" / f"{ckpt}-lora.safetensors"


def patch_trainer(base_task_cls=None):
    """Wraps `BaseTask.from_run_id` and `BaseTask.prepare_batch` (see the module docstring).  Idempotent."""
    if base_task_cls is None:
        import tasks.base as tb
        base_task_cls = tb.BaseTask
    cls = base_task_cls
    if getattr(cls, "_mts_patched", False):
        return cls
    import torch

    orig_from_run_id = cls.from_run_id.__func__
    orig_prepare = cls.prepare_batch

    def from_run_id(klass, run_id, cfg=None, ckpt="latest", basepath=None):
        trainer = orig_from_run_id(klass, run_id, cfg, ckpt, basepath)
        model = trainer.model
        if getattr(model, "lora_enabled", False) and hasattr(model.llm, "load_pretrained"):
            ckpt_name = ckpt or "latest"
            if basepath is None:      # the default of tasks/base.py:286-287 and loggers/base_logger.py:14-17
                import tasks.base as tb
                base = Path(tb.__file__).parent / "../outputs/logs"
            else:
                base = Path(basepath)
            path = lora_checkpoint_path(base, run_id, ckpt_name)
            if path.exists():
                model.llm.load_pretrained(path)
            else:
                import warnings
                warnings.warn(f"lora.enabled but {path} does not exist: the LoRA pairs keep their initial values")
        return trainer

    def prepare_batch(self, batch):
        dev_batch = orig_prepare(self, batch)
        model = getattr(self, "model", None)
        if (os.environ.get("MTS_HOST_MIRROR", "1") == "0" or model is None or model.training
                or not getattr(model, "accepts_host_mirror", False) or not isinstance(batch, dict)
                or not isinstance(dev_batch, dict) or getattr(self.device, "type", "cpu") != "cuda"):
            return dev_batch
        host = HostMirrorBatch()
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                # what the reference's loop would read back with .cpu(): same values, trainer dtype (tasks/base.py:206-208)
                host[k] = v.to(self.dtype) if v.dtype.is_floating_point and v.dtype != self.dtype else v
            else:
                host[k] = dev_batch[k]
        host.device_batch = dev_batch
        return host

    cls.from_run_id = classmethod(from_run_id)
    cls.prepare_batch = prepare_batch
    cls._mts_patched = True
    return cls


def _patch_trainer_for_dp():
    import torch
    import torch.distributed as dist
    from . import dp
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tasks.base as tb
    orig_init = tb.BaseTask.__init__

    def init(self, run_id, config, newrun=True):
        if dist.get_rank() != 0:
            config.DEBUG = True          # loggers/__init__.py:8-9 -> DebugLogger: no files, no wandb
        orig_init(self, run_id, config, newrun)
        self.train_dataloader = dp.distributed_dataloader(self.train_dataloader, seed=config.setup.seed)

    tb.BaseTask.__init__ = init
    tb.BaseTask.get_device = lambda self: torch.device("cuda", local)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m medtsllm_b200.plugin <reference>/train.py <config.toml> [...]")
    script = os.path.abspath(argv[0])
    root = os.path.dirname(script)
    sys.path.insert(0, root)
    register(root)
    patch_trainer()
    _patch_trainer_for_dp()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
