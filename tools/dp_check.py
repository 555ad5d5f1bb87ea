"""torchrun --nproc-per-node 2 tools/dp_check.py — DP equivalence on real GPUs over NCCL:
mean over ranks of per-shard gradients (what the backward's all-reduce produces) must equal the
single-process gradient of the mean loss on the concatenated batch."""
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "med-ts-llm_b200", REPO / "tests"):
    sys.path.insert(0, str(p))
import torch
import torch.distributed as dist
from _fixtures import Cfg, Dataset, config_for, load_case, materialize_llm_dir
from medtsllm_b200 import dp
from medtsllm_b200.model import MedTsLLM

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
fix = load_case("llama_seg_concat")
tmp = Path(tempfile.mkdtemp(prefix=f"dp{rank}_"))
model = MedTsLLM(Cfg(config_for(fix, materialize_llm_dir(fix, tmp / "llm"))), Dataset(fix["dataset"]))
model.load_state_dict(fix["adapters"])
model = model.to(dev).train()
x = fix["inputs"]["x_enc"].to(dev)
if x.shape[0] < 2 * world:            # every rank gets at least two windows (the fixture holds four)
    reps = -(-2 * world // x.shape[0])
    x = (x.repeat(reps, 1, 1) * torch.linspace(0.8, 1.2, reps * x.shape[0], device=dev)[:, None, None])[:2 * world].contiguous()
B = x.shape[0]
w = torch.randn(B, fix["config"]["pred_len"], generator=torch.Generator().manual_seed(1)).to(dev)


def grads(idx, sync):
    for p in model.parameters():
        p.grad = None
    import contextlib
    with (contextlib.nullcontext() if sync else dp.suspended()):      # single-process reference: no all-reduce
        out = model({"x_enc": x[idx]})
        (out * w[idx]).sum().backward()
    return {k: p.grad.clone() for k, p in model.named_parameters()}


shard = list(dp.shard_batch(B, rank, world))
g_dp = grads(shard, sync=True)                       # all-reduced mean over ranks
g_full = grads(list(range(B)), sync=False)           # sum over the full batch, locally
worst = 0.0
for k in g_dp:
    ref = g_full[k] / world
    err = ((g_dp[k] - ref).norm() / ref.norm().clamp_min(1e-20)).item()
    if k != "reprogramming_layer.key_projection.bias":
        worst = max(worst, err)
t = torch.tensor([worst], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"[dp_check] world={world} shard sizes={[len(list(dp.shard_batch(B, r, world))) for r in range(world)]} "
          f"worst rel-L2 (DP mean vs full-batch/world) = {t.item():.2e}")
    assert t.item() < 2e-2, t.item()      # bf16 GEMMs over different batch splits: summation-order noise

# training-step graphs under data parallelism (GPT-2-sized backbone -> "auto" uses them): the same shard three times
# (kernel by kernel, capture, replay; the all-reduce then runs after the backward graph) against the overlapped
# kernel-by-kernel path
model.use_train_graph = "1"
for _ in range(3):
    g_graph = grads(shard, sync=True)
assert model._train_graph is not None and model._train_graph.entry is not None and model._train_graph.entry["bwd"] is not None
model.use_train_graph = "0"
g_eager = grads(shard, sync=True)
worst2 = 0.0
for k in g_eager:
    if k != "reprogramming_layer.key_projection.bias":
        worst2 = max(worst2, ((g_graph[k] - g_eager[k]).norm() / g_eager[k].norm().clamp_min(1e-20)).item())
t = torch.tensor([worst2], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"[dp_check] graph replay + all-reduce vs overlapped kernel-by-kernel path: worst rel-L2 = {t.item():.2e}")
    assert t.item() < 1e-4, t.item()      # same kernels; only the atomically reduced conv gradient differs in its last bits
dist.destroy_process_group()
