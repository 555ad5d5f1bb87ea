#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
grep -E "passed|failed|FAILED|Error|lora shared" gpurun_out/pytest_gpu.log | tail -20 >> gpurun_out/round.log
run bench_vent 900 python bench.py --steps 10 --warmup 3 --workload ventilator_llama2_7b --no-cpu-baseline --no-ref-gpu
tail -1 gpurun_out/bench_vent.log | cut -c1-3000 >> gpurun_out/round.log
cat gpurun_out/round.log
