"""Data parallelism for the hot path: one process per GPU, batch sharded across ranks, frozen backbone
replicated, ONE exchange step per training iteration — an all-reduce (mean) of the trainable adapter
gradients only (the reference itself has no distributed code at all).

The forward has no collective.  In the backward the head / down-sample gradients are complete before
the dgrad through the 32 frozen blocks starts, so their bucket is all-reduced asynchronously (NCCL
runs on its own stream) underneath the backbone backward; the mapping / reprogramming / patch-embed
gradients only exist after the backbone dgrad and go in a second bucket at the end.

Two things keep the exposed (non-overlapped) part of that second exchange small:
  * gradients are produced straight into slices of ONE pre-allocated flat fp32 buffer per bucket
    (`GradArena`), so the all-reduce runs in place: no `torch.cat`, no copy back, and the 1/world
    scaling is NCCL's own `ReduceOp.AVG`;
  * the mapping-layer weight gradient (S x V fp32: 131 MB for Llama-2-7B, 206 MB for GPT-2-medium — by far
    the largest late tensor) is never exchanged.  It is `dSource . E^T` with the frozen embedding table E
    identical on every rank, and the all-reduce is linear, so the ranks average `dSource` (S x D fp32:
    16.8 MB / 4.2 MB) and each computes the GEMM on the averaged operand (train._backward_chain).

torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is the
plumbing; nothing here touches the kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


_suspended = False


def is_active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def comm_enabled() -> bool:
    """is_active() and not inside `suspended()`: whether the gradient exchanges are actually issued."""
    return is_active() and not _suspended


class suspended:
    """Context manager: run training steps WITHOUT the gradient exchange (every rank keeps its local gradients; the
    choice between the captured-graph and the kernel-by-kernel step is unaffected).  Measurement only — bench.py times the same step with and without the all-reduces to report the exposed
    communication cost of data parallelism (`train_step.dp_efficiency`)."""

    def __enter__(self):
        global _suspended
        self._prev, _suspended = _suspended, True
        return self

    def __exit__(self, *exc):
        global _suspended
        _suspended = self._prev
        return False


def _avg_supported(group=None) -> bool:
    try:
        return dist.get_backend(group) == "nccl"
    except Exception:  # noqa: BLE001
        return False


def all_reduce_mean_(flat: torch.Tensor, group=None, async_op: bool = True):
    """In-place mean over ranks.  NCCL averages inside the collective; gloo (CPU tests) sums and the caller's
    `finish` divides.  Returns (work handle, needs_div)."""
    if _avg_supported(group):
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op), False
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op), True


class GradArena:
    """One flat fp32 buffer holding a fixed, ordered set of gradient tensors; `view(name)` is the slice a kernel
    writes its result into.  The all-reduce runs in place on the whole buffer (or on a prefix / suffix range)."""

    def __init__(self, named_shapes, device):
        self.offsets = {}
        off = 0
        for name, shape in named_shapes:
            n = 1
            for s in shape:
                n *= int(s)
            self.offsets[name] = (off, n, tuple(shape))
            off += (n + 3) // 4 * 4            # 16-byte aligned slices (vectorised epilogue stores)
        self.flat = torch.empty(off, device=device, dtype=torch.float32)
        self.handle = None
        self._needs_div = False
        self.group = None

    def view(self, name: str) -> torch.Tensor:
        off, n, shape = self.offsets[name]
        return self.flat[off:off + n].view(shape)

    def __contains__(self, name):
        return name in self.offsets

    def launch(self, group=None):
        """Start the (async) in-place mean all-reduce of the whole arena.  The collective is enqueued after the
        work already on the current stream (torch's ProcessGroupNCCL waits on the current stream's tail)."""
        if self.flat.numel() == 0 or not comm_enabled():
            return self
        self.group = group
        self.handle, self._needs_div = all_reduce_mean_(self.flat, group, async_op=True)
        return self

    def finish(self):
        if self.handle is None:
            return
        self.handle.wait()
        if self._needs_div:
            self.flat.div_(dist.get_world_size(self.group))
        self.handle = None


class GradBucket:
    """Flattens a list of gradient tensors into one contiguous buffer, all-reduces it (async) and
    scatters the mean back in place.  Kept for gradients that cannot be produced into a `GradArena`
    (LoRA pairs, GPT4TS); the MedTsLLM adapters use arenas."""

    def __init__(self, tensors, group=None):
        self.tensors = [t for t in tensors if t is not None]
        self.group = group
        self.handle = None
        self.flat = None
        self._needs_div = False

    def launch(self):
        if not self.tensors or not comm_enabled():
            return self
        self.flat = torch.cat([t.reshape(-1) for t in self.tensors])
        self.handle, self._needs_div = all_reduce_mean_(self.flat, self.group, async_op=True)
        return self

    def finish(self):
        if self.handle is None:
            return
        self.handle.wait()
        if self._needs_div:
            self.flat.div_(dist.get_world_size(self.group))
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


class EpochShuffledLoader:
    """A DataLoader over a DistributedSampler whose `__iter__` advances the sampler's epoch.  The reference's
    train loops just iterate `self.train_dataloader` once per epoch (tasks/forecasting.py:15-19) and its own
    `DataLoader(shuffle=True)` reshuffles on every pass; a bare DistributedSampler would replay the same
    permutation (and the same shard assignment) every epoch unless somebody calls `set_epoch`."""

    def __init__(self, loader, sampler):
        self._loader, self.sampler, self._epoch = loader, sampler, 0

    def __iter__(self):
        self.sampler.set_epoch(self._epoch)
        self._epoch += 1
        return iter(self._loader)

    def __len__(self):
        return len(self._loader)

    def __getattr__(self, name):                 # batch_size, dataset, collate_fn, num_workers, pin_memory, ...
        return getattr(self._loader, name)


def distributed_dataloader(loader, rank: int | None = None, world: int | None = None, seed: int = 0):
    """Rebuilds a reference DataLoader (tasks/base.py:175-182: shuffle=True, no sampler — every rank
    would see the same batches) around a DistributedSampler, keeping batch size, collate_fn, workers and
    pin_memory; every pass over the result is a fresh epoch (new permutation, as `shuffle=True` gives the
    reference).  Used by the launcher to swap `trainer.train_dataloader` without editing tasks/*."""
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    sampler = DistributedSampler(loader.dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed,
                                 drop_last=False)
    inner = DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=sampler,
                       collate_fn=loader.collate_fn, num_workers=loader.num_workers, pin_memory=loader.pin_memory)
    return EpochShuffledLoader(inner, sampler)
