"""Two launches of the TF32 GEMM of the evaluation parity mode at the BIDMC qkv shape (plain kind::tf32, then 3xTF32),
for `ncu --set full` (tools/gpu_round2.sh)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "med-ts-llm_b200"))
import torch
from medtsllm_b200 import ops
dev = torch.device("cuda:0")
m, n, k = 2176, 12288, 4096
a = torch.randn(m, k, device=dev)
b = torch.randn(n, k, device=dev) * 0.02
d = torch.empty(m, n, device=dev)
ah, al = ops.split_tf32(a)
bh, bl = ops.split_tf32(b)
torch.cuda.synchronize()
ops.gemm(ah, bh, d, m=m, n=n, k=k)
ops.gemm(ah, bh, d, m=m, n=n, k=k, a_lo=al, b_lo=bl)
torch.cuda.synchronize()
print("ok")
