"""Drop-in registration into the reference (flixpar/med-ts-llm) — the plug-in seam of SURVEY.md §8b.

The reference builds its model with `model_lookup[config.model](config, train_dataset)`
(tasks/base.py:81-85, models/__init__.py:10-18).  `register()` points the keys "medtsllm" and
"timellm" at `medtsllm_b200.MedTsLLM`; `train.py`, `tasks/*`, `datasets/*` and `loggers/*` then run
unchanged:

    python -m medtsllm_b200.plugin /path/to/med-ts-llm/train.py configs/datasets/bidmc.toml
    torchrun --nproc-per-node 8 -m medtsllm_b200.plugin /path/to/med-ts-llm/train.py cfg.toml   # DP

With WORLD_SIZE > 1 the launcher also (1) initialises NCCL, (2) swaps each trainer's
`train_dataloader` for a DistributedSampler-backed one right after BaseTask.__init__ (the reference has
`shuffle=True` and no sampler, tasks/base.py:175-182), and (3) keeps real logging on rank 0 only.
Gradient all-reduce happens inside the model's backward (medtsllm_b200/dp.py).
"""
from __future__ import annotations

import os
import runpy
import sys


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
    _patch_trainer_for_dp()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
