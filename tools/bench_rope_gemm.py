"""qkv projection with the fused RoPE epilogue (mts_gemm, MTS_EPI_ROPE_QK) at the shared-prefix row counts of the
Ventilator / BIDMC configs, 20 launches replayed from a CUDA graph (the form tools/bench_gemm.py --graph uses)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "med-ts-llm_b200"))
import torch
from medtsllm_b200 import _lib, ops

dev = torch.device("cuda:0")


def tabs(L, hd):
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2, device=dev).float() / hd))
    f = torch.outer(torch.arange(L, device=dev).float(), inv)
    return f.cos().contiguous(), f.sin().contiguous()


for (name, m, n, k, Lc, Ls) in (("vent_qkv_rope", 800, 12288, 4096, 128, 42), ("sp_qkv_rope", 2176, 12288, 4096, 128, 64)):
    A = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(3)]
    B = [(torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16) for _ in range(3)]
    D = torch.zeros(m, n, device=dev, dtype=torch.bfloat16)
    rope = tabs(Lc + Ls, 128)

    def run(i):
        ops.gemm(A[i % 3], B[i % 3], D, m=m, n=n, k=k, epilogue=_lib.EPI_ROPE_QK, rope=rope, rope_L=Ls, rope_hd=128,
                 rope_cols=8192, rope_prefix=Lc)

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(20):
            run(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, round(ms * 1000, 1), "us", round(2.0 * m * n * k / ms / 1e9, 1), "TF/s", flush=True)
