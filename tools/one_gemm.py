#!/usr/bin/env python
"""A few launches of one small-M GEMM of the PSM / GPT-2-medium layer on the single-CTA kernel (MTS_GEMM_DBG=1 prints the
kernel's clock stamps):  python tools/one_gemm.py [o|proj|qkv|fc]"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "med-ts-llm_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402
from medtsllm_b200 import _lib, ops  # noqa: E402

SHAPES = {"o": (896, 1024, 1024, 1), "proj": (896, 1024, 4096, 1), "qkv": (896, 3072, 1024, 0), "fc": (896, 4096, 1024, 2)}
m, n, k, epi = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "o"]
dev = torch.device("cuda", 0)
_lib.set_option("gemm_force", 1)
a = (torch.randn(m, k, device=dev) * 0.1).to(torch.bfloat16)
ws = [(torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16) for _ in range(40)]     # rotate weights: HBM, not L2
bias = torch.zeros(n, device=dev)
d = torch.zeros(m, n, device=dev) if epi == 1 else torch.empty(m, n, device=dev, dtype=torch.bfloat16)
for i in range(6):
    ops.gemm(a, ws[i], d, m=m, n=n, k=k, bias=bias, bias_axis=1, epilogue=epi)
torch.cuda.synchronize()
print("ok")
