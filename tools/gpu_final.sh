#!/bin/bash
# GPU tests + smoke + the headline bench line with its HBM table.
cd "$(dirname "$0")/.."
bash tools/gpu_tests.sh 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 ${1:---no-cpu-baseline --no-ref-gpu} > gpurun_out/bench_main.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_main.log").read().strip().splitlines()[-1])
print("bidmc fwd", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "train", d["train_step"]["value"], d["train_step"]["ms_per_step"],
      "clk", d["clocks"]["sm_mhz"], "gemm", d["roofline"]["achieved"], d["roofline"]["frac"])
for r in d["hbm_roofline"]:
    print("  ", r)
PY
