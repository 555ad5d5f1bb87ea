#!/bin/bash
# One GPU-box visit: model parity tests, smoke, bench (mini first, then the headline workload).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest model" > gpurun_out/round.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s > gpurun_out/test_model.log 2>&1; echo "exit $?" >> gpurun_out/round.log
grep -E "parity|passed|failed|Error|error" gpurun_out/test_model.log | tail -20 >> gpurun_out/round.log
echo "=== smoke" >> gpurun_out/round.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/round.log
tail -5 gpurun_out/smoke.log >> gpurun_out/round.log
echo "=== bench mini" >> gpurun_out/round.log
timeout 300 python bench.py --workload mini_llama --steps 5 --warmup 3 > gpurun_out/bench_mini.log 2>&1; echo "exit $?" >> gpurun_out/round.log
tail -3 gpurun_out/bench_mini.log >> gpurun_out/round.log
echo "=== bench headline" >> gpurun_out/round.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_main.log 2>&1; echo "exit $?" >> gpurun_out/round.log
tail -3 gpurun_out/bench_main.log >> gpurun_out/round.log
cat gpurun_out/round.log
