"""Micro-benchmark of mts_gemm at the backbone shapes (CUDA events, L2-sized rotation of operands)."""
import json, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "med-ts-llm_b200"))
import torch
from medtsllm_b200 import ops

dev = torch.device("cuda:0")
shapes = [  # (name, m, n, k, epilogue, block_n)
    # shared-prefix row counts of the BIDMC step (128 + 32*64 rows)
    ("sp_qkv", 2176, 12288, 4096, 0, 0),
    ("sp_o_resid", 2176, 4096, 4096, 1, 0),
    ("sp_gateup_swiglu", 2176, 22016, 4096, 3, 256),
    ("sp_down_resid", 2176, 4096, 11008, 1, 0),
    ("llama_qkv", 6144, 12288, 4096, 0, 0),
    ("llama_o_resid", 6144, 4096, 4096, 1, 0),
    ("llama_gateup_swiglu", 6144, 22016, 4096, 3, 256),
    ("llama_down_resid", 6144, 4096, 11008, 1, 0),
    ("llama_qkv_bn128", 6144, 12288, 4096, 0, 128),
    ("gpt2m_qkv", 8960, 3072, 1024, 0, 0),
    ("gpt2m_fc_gelu", 8960, 4096, 1024, 2, 0),
    ("gpt2m_proj_resid", 8960, 1024, 4096, 1, 0),
    ("mapping", 1024, 4096, 32000, 0, 0),
    # small-M regime (shared-prefix row counts of Ventilator: 128 + 16*42 = 800, PSM: 128 + 64*12 = 896)
    ("vent_qkv", 800, 12288, 4096, 0, 0),
    ("vent_o_resid", 800, 4096, 4096, 1, 0),
    ("vent_gateup_swiglu", 800, 22016, 4096, 3, 256),
    ("vent_down_resid", 800, 4096, 11008, 1, 0),
    ("psm_qkv", 896, 3072, 1024, 0, 0),
    ("psm_o_resid", 896, 1024, 1024, 1, 0),
    ("psm_fc_gelu", 896, 4096, 1024, 2, 0),
    ("psm_proj_resid", 896, 1024, 4096, 1, 0),
]
from medtsllm_b200 import _lib
res = []
force_modes = [0, 1, 2] if "--force-sweep" in sys.argv else [0]      # auto / single-CTA kernel / CTA-pair kernel
if "--shared-prefix-only" in sys.argv:
    shapes = shapes[:4]
if "--small-m" in sys.argv:
    shapes = shapes[-8:]
sk_modes = [0, 1, 2] if "--streamk-sweep" in sys.argv else [None]
if "--splitk-sweep" in sys.argv:     # even split-K (mode 3) at forced tile widths, against the auto schedule
    shapes = [s for s in shapes if s[0] in ("psm_o_resid", "psm_proj_resid", "vent_o_resid")]
    shapes = [(nm, m, n, k, e, bn) for (nm, m, n, k, e, _) in shapes for bn in (0, 64, 128, 256)]
    sk_modes = [0, 3]
ks_modes = [None]
if "--ksplit-sweep" in sys.argv:     # cluster split-K at forced (tile width, cluster size) against the plain schedules
    shapes = [s for s in shapes if s[0].startswith("psm_") and s[4] != 3] + [s for s in shapes if s[0] == "vent_o_resid"]
    shapes = [(nm, m, n, k, e, bn) for (nm, m, n, k, e, _) in shapes for bn in (0, 64, 128, 256)]
    ks_modes = [0, 2, 4, -1]
for (name, m, n, k, epi, bn), force, skm, ksm in ((sh, f, sk, ks) for sh in shapes for f in force_modes for sk in sk_modes
                                                 for ks in ks_modes):
    if ksm is not None:
        if (bn == 0) != (ksm in (0, -1)) and not (bn != 0 and ksm == 0):
            continue          # auto width only with off / auto; forced widths with off / 2 / 4
        _lib.set_option("gemm_ksplit", ksm)
    _lib.set_option("gemm_force", force)
    if skm is not None:
        ops.set_streamk(skm)
    nrot = 3
    A = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nrot)]
    B = [(torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16) for _ in range(nrot)]
    n_out = n // 2 if epi == 3 else n
    D = torch.zeros(m, n_out, device=dev, dtype=torch.float32 if epi == 1 else torch.bfloat16)
    bias = torch.zeros(n, device=dev) if epi == 2 else None
    def run(i):
        ops.gemm(A[i % nrot], B[i % nrot], D, m=m, n=n, k=k, epilogue=epi, block_n=bn,
                 bias=bias, bias_axis=1 if bias is not None else 0)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    use_graph = "--graph" in sys.argv      # launches replayed from a CUDA graph: no host launch-rate floor (~15 us from Python)
    def timed(fn):
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(iters): fn(i)
            g.replay(); torch.cuda.synchronize()
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        else:
            e0.record()
            for i in range(iters): fn(i)
            e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    ms = timed(run)
    # cuBLAS yardstick (library GEMM, plain store)
    C = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for i in range(3): torch.matmul(A[i % nrot], B[i % nrot].t(), out=C)
    torch.cuda.synchronize()
    ms_cublas = timed(lambda i: torch.matmul(A[i % nrot], B[i % nrot].t(), out=C))
    fl = 2.0 * m * n * k
    r = dict(name=name, force=force, streamk=skm, ksplit=ksm, m=m, n=n, k=k, epi=epi, bn=bn, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1),
             cublas_ms=round(ms_cublas, 4), cublas_tflops=round(fl / ms_cublas / 1e9, 1))
    print(json.dumps(r), flush=True)
    res.append(r)
    del A, B, D, C
Path(REPO / "gpurun_out").mkdir(exist_ok=True)
(REPO / "gpurun_out" / "bench_gemm.json").write_text(json.dumps(res, indent=1))
