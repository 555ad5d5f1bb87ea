#!/bin/bash
# Full ncu capture of the attention backward kernel(s) of one training step + the train-graph test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "training_step_graphs or full_size" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd --profile-from-start off -s 0 -c 2 -f -o gpurun_out/prof_attnbwd python bench.py --profile-step train --no-cpu-baseline --no-train --no-ref-gpu > gpurun_out/ncu_full_attnbwd.log 2>&1; echo "ncu exit $?"
