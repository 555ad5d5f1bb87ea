"""Data parallelism for the hot path: one process per GPU, batch sharded across ranks, frozen backbone
replicated, ONE exchange step per training iteration — an all-reduce (mean) of the trainable adapter
gradients only (the reference itself has no distributed code at all).

The forward has no collective.  In the backward the head / down-sample gradients are complete before
the dgrad through the 32 frozen blocks starts, so their bucket is all-reduced asynchronously (NCCL
runs on its own stream) underneath the backbone backward; the mapping / reprogramming / patch-embed
gradients only exist after the backbone dgrad and go in a second bucket at the end.  Payload with the
shipped configs: <= ~115 M fp32 values.

torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is the
plumbing; nothing here touches the kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class GradBucket:
    """Flattens a list of gradient tensors into one contiguous buffer, all-reduces it (async) and
    scatters the mean back in place."""

    def __init__(self, tensors, group=None):
        self.tensors = [t for t in tensors if t is not None]
        self.group = group
        self.handle = None
        self.flat = None

    def launch(self):
        if not self.tensors or not is_active():
            return self
        self.flat = torch.cat([t.reshape(-1) for t in self.tensors])
        self.handle = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return self

    def finish(self):
        if self.handle is None:
            return
        self.handle.wait()
        world = dist.get_world_size(self.group)
        self.flat.div_(world)
        off = 0
        for t in self.tensors:
            n = t.numel()
            t.copy_(self.flat[off:off + n].view_as(t))
            off += n
        self.handle = None
        self.flat = None


def shard_batch(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of a batch (the first n % world ranks get one extra item)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def distributed_dataloader(loader, rank: int | None = None, world: int | None = None, seed: int = 0):
    """Rebuilds a reference DataLoader (tasks/base.py:175-182: shuffle=True, no sampler — every rank
    would see the same batches) around a DistributedSampler, keeping batch size, collate_fn, workers and
    pin_memory.  Used by the launcher to swap `trainer.train_dataloader` without editing tasks/*."""
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    sampler = DistributedSampler(loader.dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed,
                                 drop_last=False)
    return DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=sampler,
                      collate_fn=loader.collate_fn, num_workers=loader.num_workers, pin_memory=loader.pin_memory)
