#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run test_attn 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attn or rope"
grep -E "passed|failed|Error|assert" gpurun_out/test_attn.log | tail -8 >> gpurun_out/round.log
run test_model 1500 python -m pytest tests/test_model_gpu.py -m gpu -q
grep -E "passed|failed|Error" gpurun_out/test_model.log | tail -8 >> gpurun_out/round.log
run bench_main 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
tail -1 gpurun_out/bench_main.log | cut -c1-1500 >> gpurun_out/round.log
cat gpurun_out/round.log
