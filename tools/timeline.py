#!/usr/bin/env python
"""Kernel timeline of one step (nsys is not in the image: CUPTI through torch.profiler instead).

  python tools/timeline.py [--workload bidmc_llama2_7b] [--mode fwd|train] [--out profiles/r02_timeline_fwd.md]

Runs warm steps of the workload, records ONE step's kernels (graph replay included) with start / end timestamps and
reports: step span, sum of kernel time, idle time between consecutive kernels (total and by the kernel that follows the
gap), overlap (PDL prologues), and the per-kernel-name totals.  Answers "where does the part of ms_per_step that no
kernel accounts for go"."""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "med-ts-llm_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import torch  # noqa: E402


def short(name: str) -> str:
    name = name.replace("void ", "").replace("mts::", "")
    return name.split("(")[0][:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="bidmc_llama2_7b")
    ap.add_argument("--mode", default="fwd", choices=["fwd", "train"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="evaluation precision mode (fwd only): tf32 / fp32 = the parity modes of precise.py")
    args = ap.parse_args()
    from medtsllm_b200.backbone import KernelBackbone
    from medtsllm_b200.model import MedTsLLM
    from medtsllm_b200.synthetic import (WORKLOADS, AttrDict, FixedLengthTokenizer, SyntheticDataset,
                                         experiment_config, make_inputs)
    dev = torch.device("cuda", 0)
    w = WORKLOADS[args.workload]
    bb = KernelBackbone.random_init(w.backbone, dev, seed=0, precision=args.precision)
    torch.manual_seed(0)
    model = MedTsLLM(AttrDict(experiment_config(w)), SyntheticDataset(w), backbone=bb,
                     tokenizer=FixedLengthTokenizer(w.backbone.vocab, w.prompt_len)).to(dev, torch.float32)
    if args.no_graph:
        model.use_cuda_graph = False
        model.use_train_graph = "0"
    x = make_inputs(w)["x_enc"].to(dev)
    if args.mode == "fwd":
        model.eval()

        def step():
            with torch.no_grad():
                model({"x_enc": x})
    else:
        model.train()
        opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)

        def step():
            loss = model({"x_enc": x}).float().pow(2).mean()
            loss.backward()
            opt.step()
            opt.zero_grad()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        prof.export_chrome_trace(f.name)
        trace = json.load(open(f.name))
    ev = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    if not ev:
        raise SystemExit("no device events recorded")
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    span = t1 - t0
    busy = 0.0
    cur_end = t0
    gaps_after = defaultdict(float)       # idle time preceding a kernel, keyed by that kernel's name
    gap_count = defaultdict(int)
    overlap = 0.0
    per_name = defaultdict(lambda: [0, 0.0])
    for e in ev:
        s, d = e["ts"], e["dur"]
        nm = short(e["name"])
        per_name[nm][0] += 1
        per_name[nm][1] += d
        if s > cur_end:
            gaps_after[nm] += s - cur_end
            gap_count[nm] += 1
        else:
            overlap += min(cur_end, s + d) - s
        busy += max(0.0, s + d - max(cur_end, s))
        cur_end = max(cur_end, s + d)
    idle = span - busy
    lines = [f"# Kernel timeline — {args.workload}, {args.mode} step, precision {args.precision} ({'kernel by kernel' if args.no_graph else 'graph replay'})", "",
             f"CUPTI (torch.profiler), one warm step on one B200; {len(ev)} device activities.", "",
             f"* step span (first kernel start -> last kernel end): **{span / 1e3:.3f} ms**",
             f"* device busy (union of kernel intervals): {busy / 1e3:.3f} ms = {100 * busy / span:.1f} % of the span",
             f"* idle between kernels: {idle / 1e3:.3f} ms = {100 * idle / span:.1f} %",
             f"* sum of kernel durations: {sum(v[1] for v in per_name.values()) / 1e3:.3f} ms "
             f"(overlapped prologue time under programmatic dependent launch: {overlap / 1e3:.3f} ms)", "",
             "| kernel | launches | total ms | share of span | avg us | idle before it: total us (avg us) |", "|---|---:|---:|---:|---:|---:|"]
    for nm, (n, tot) in sorted(per_name.items(), key=lambda kv: -kv[1][1]):
        g = gaps_after.get(nm, 0.0)
        lines.append(f"| `{nm}` | {n} | {tot / 1e3:.3f} | {100 * tot / span:.1f} % | {tot / n:.1f} | "
                     f"{g:.0f} ({g / max(gap_count.get(nm, 0), 1):.1f}) |")
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(text)


if __name__ == "__main__":
    main()
