#!/usr/bin/env python
"""A few launches of the attention forward at one BASELINE shape (for ncu):  python tools/one_attn.py [bidmc|psm|ludb|vent]"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "med-ts-llm_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402
from medtsllm_b200 import ops  # noqa: E402

SHAPES = {"bidmc": (32, 128, 64, 32, 128), "ludb": (16, 128, 128, 32, 128), "vent": (16, 128, 42, 32, 128),
          "psm": (64, 128, 12, 16, 64)}
Bp, Lc, Ls, H, hd = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "bidmc"]
dev = torch.device("cuda", 0)
qkv = (torch.randn(Lc + Bp * Ls, 3 * H * hd, device=dev) * 0.5).to(torch.bfloat16)
for _ in range(4):
    out, lse = ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, want_lse=True)
if "--bwd" in sys.argv:
    dout = (torch.randn(Bp * Ls, H * hd, device=dev) * 0.5).to(torch.bfloat16)
    lse_own = ops.lse_own_view(lse, Bp, Lc, Ls, H)
    for _ in range(3):
        ops.attn_causal_shared_bwd(qkv, out[Lc:], dout, lse_own, Bp, Lc, Ls, H, hd)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
