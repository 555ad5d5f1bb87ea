#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run test_model_train 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "training"
grep -E "parity|passed|failed|Error|error" gpurun_out/test_model_train.log | tail -20 >> gpurun_out/round.log
run bench_mini 600 python bench.py --workload mini_llama --steps 6 --warmup 3 --no-cpu-baseline
tail -2 gpurun_out/bench_mini.log >> gpurun_out/round.log
run bench_main 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
tail -2 gpurun_out/bench_main.log >> gpurun_out/round.log
cat gpurun_out/round.log
