"""Correctness of the CTA-pair GEMM (cta_group::2) against torch, shape by shape (run under `timeout`)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "med-ts-llm_b200"))
import torch
from medtsllm_b200 import _lib, ops

dev = torch.device("cuda:0")
_lib.set_option("gemm_2cta", 1)
g = torch.Generator().manual_seed(0)
ok = True
for (m, n, k, epi) in [(256, 256, 64, 0), (256, 256, 256, 0), (512, 512, 512, 0), (384, 768, 320, 0), (300, 520, 200, 0),
                       (6144, 4096, 4096, 1), (2048, 1024, 1024, 2), (1000, 1536, 512, 3), (6144, 12288, 4096, 0)]:
    a = (torch.randn(m, k, generator=g) * 0.5).to(dev, torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.5).to(dev, torch.bfloat16)
    ref = a.float() @ b.float().t()
    if epi == 0:
        d = torch.full((m, n), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.gemm(a, b, d, m=m, n=n, k=k, block_n=256)
    elif epi == 1:
        c = torch.randn(m, n, generator=g).to(dev)
        d = c.clone()
        ops.gemm(a, b, d, m=m, n=n, k=k, block_n=256, epilogue=1)
        ref = ref + c
    elif epi == 2:
        bias = torch.randn(n, generator=g).to(dev)
        d = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, b, d, m=m, n=n, k=k, block_n=256, epilogue=2, bias=bias, bias_axis=1)
        x = ref + bias
        ref = 0.5 * x * (1 + torch.tanh(0.7978845608 * (x + 0.044715 * x ** 3)))
    else:
        d = torch.empty(m, n // 2, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, b, d, m=m, n=n, k=k, epilogue=3)
        r4 = ref.view(m, n // 256, 2, 128)
        ref = (torch.nn.functional.silu(r4[:, :, 0]) * r4[:, :, 1]).reshape(m, n // 2)
    torch.cuda.synchronize()
    err = ((d.float() - ref).norm() / ref.norm()).item()
    print(f"2cta m={m} n={n} k={k} epi={epi}: rel-L2 {err:.2e}", flush=True)
    ok &= err < 5e-3
print("2CTA_OK" if ok else "2CTA_FAIL")
