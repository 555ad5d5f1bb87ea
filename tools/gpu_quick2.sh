#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -20 >> gpurun_out/round.log
for wl in psm_gpt2_medium bidmc_llama2_7b; do
run bench_$wl 900 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-ref-gpu
python - <<PY >> gpurun_out/round.log
import json
d=json.loads(open("gpurun_out/bench_$wl.log").read().strip().splitlines()[-1])
print("$wl", "fwd", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "train", d["train_step"])
PY
done
cat gpurun_out/round.log
