"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic: gradient bucketing/averaging, batch
sharding, and the DistributedSampler swap of the reference's DataLoader."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from medtsllm_b200 import dp
        assert dp.is_active()
        g = torch.Generator().manual_seed(100 + rank)
        grads = [torch.randn(7, 5, generator=g), None, torch.randn(3, generator=g), torch.randn(2, 2, 2, generator=g)]
        mine = [t.clone() if t is not None else None for t in grads]
        early = dp.GradBucket(grads[:2]).launch()
        late = dp.GradBucket(grads[2:]).launch()
        early.finish(); late.finish()
        # expected: mean over ranks of the per-rank tensors
        for i, t in enumerate(grads):
            if t is None:
                continue
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, mine[i])
            torch.testing.assert_close(t, torch.stack(parts).mean(0), rtol=1e-6, atol=1e-6)
        # gradient arena: kernels write straight into slices of one flat buffer, all-reduced in place (no cat / copy)
        arena = dp.GradArena([("w", (4, 3)), ("b", (5,)), ("c", (2, 2))], "cpu")
        assert all(arena.offsets[k][0] % 4 == 0 for k in arena.offsets)          # 16-byte aligned slices
        local = {k: torch.randn(arena.view(k).shape, generator=g) for k in ("w", "b", "c")}
        for k, t in local.items():
            arena.view(k).copy_(t)
        arena.launch().finish()
        for k, t in local.items():
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            torch.testing.assert_close(arena.view(k), torch.stack(parts).mean(0), rtol=1e-6, atol=1e-6)
        # batch sharding covers every item exactly once
        shards = [list(dp.shard_batch(11, r, world)) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(11)) and abs(len(shards[0]) - len(shards[1])) <= 1
        # DataLoader swap: disjoint halves of the dataset, same batch size / collate
        ds = torch.utils.data.TensorDataset(torch.arange(20))
        loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=True)
        dl = dp.distributed_dataloader(loader, seed=0)
        seen = torch.cat([b[0] for b in dl])
        all_seen = [torch.empty_like(seen) for _ in range(world)]
        dist.all_gather(all_seen, seen)
        assert sorted(torch.cat(all_seen).tolist()) == list(range(20))
        assert dl.batch_size == 4 and len(dl) == 3
        # every pass is a new epoch: a different permutation (the reference's shuffle=True reshuffles per epoch,
        # tasks/base.py:175-182), still a partition of the dataset across ranks
        seen2 = torch.cat([b[0] for b in dl])
        assert not torch.equal(seen, seen2)
        all_seen2 = [torch.empty_like(seen2) for _ in range(world)]
        dist.all_gather(all_seen2, seen2)
        assert sorted(torch.cat(all_seen2).tolist()) == list(range(20))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_dp_host_logic_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_bucket_is_noop_without_process_group():
    from medtsllm_b200 import dp
    t = torch.ones(3)
    b = dp.GradBucket([t]).launch()
    b.finish()
    assert torch.equal(t, torch.ones(3)) and not dp.is_active()
