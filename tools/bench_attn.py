#!/usr/bin/env python
"""Isolated timing of the attention kernels on the shared-prefix layout at the BASELINE shapes (20 launches captured in a
CUDA graph, device time only):  python tools/bench_attn.py"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "med-ts-llm_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402
from medtsllm_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


SHAPES = [("bidmc", 32, 128, 64, 32, 128), ("ludb", 16, 128, 128, 32, 128), ("vent", 16, 128, 42, 32, 128),
          ("psm", 64, 128, 12, 16, 64), ("bidmc per-sample", 32, 0, 192, 32, 128)]
if "--hd64" in sys.argv:      # GPT-2 family shapes (head dim 64)
    SHAPES = [("psm", 64, 128, 12, 16, 64), ("psm per-sample", 64, 0, 140, 16, 64), ("gpt4ts etth1", 8, 0, 192, 12, 64),
              ("gpt2-medium bidmc", 32, 128, 64, 16, 64), ("gpt2-medium ludb", 16, 128, 128, 16, 64), ("gpt2 small", 16, 64, 24, 12, 64)]
for name, Bp, Lc, Ls, H, hd in SHAPES:
    M = Lc + Bp * Ls
    D = H * hd
    gen = torch.Generator(device=dev).manual_seed(0)
    qkv = (torch.randn(M, 3 * D, device=dev, generator=gen) * 0.5).to(torch.bfloat16)
    dout = (torch.randn(M, D, device=dev, generator=gen) * 0.5).to(torch.bfloat16)
    flops = 4.0 * H * hd * (Lc * Lc / 2 + Bp * Ls * (Lc + Ls / 2))
    # Llama shapes (head dim 128): the backward rotates dQ / dK back (RoPE), as in the training step
    Ltab = Lc + Ls
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=dev, dtype=torch.float32) / hd))
    ang = torch.arange(Ltab, device=dev, dtype=torch.float32)[:, None] * inv[None, :]
    rope = (ang.cos().contiguous(), ang.sin().contiguous()) if hd == 128 else None
    if Lc:
        out, lse = ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, want_lse=True)
        t_f = timed(lambda: ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, out=out, want_lse=True))
        own = slice(Lc, None)
        lse_own = ops.lse_own_view(lse, Bp, Lc, Ls, H)
        t_b = timed(lambda: ops.attn_causal_shared_bwd(qkv, out[own], dout[own], lse_own, Bp, Lc, Ls, H, hd, rope=rope))
    else:
        out, lse = ops.attn_causal(qkv, Bp, Ls, H, hd, want_lse=True)
        t_f = timed(lambda: ops.attn_causal(qkv, Bp, Ls, H, hd, out=out, want_lse=True))
        t_b = timed(lambda: ops.attn_causal_bwd(qkv, out, dout, lse, Bp, Ls, H, hd, rope=rope, pre_roped=True))
    print(f"[attn] {name:18s} Bp={Bp} Lc={Lc} Ls={Ls} H={H} hd={hd}: fwd {t_f:7.1f} us ({flops / t_f / 1e6:6.1f} TFLOP/s)   "
          f"bwd {t_b:7.1f} us ({2.5 * flops / t_b / 1e6:6.1f} TFLOP/s)")
