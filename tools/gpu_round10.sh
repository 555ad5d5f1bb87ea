#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run test_2cta 120 python tools/test_2cta.py
tail -4 gpurun_out/test_2cta.log >> gpurun_out/round.log
run pytest_gpu 1500 python -m pytest tests -m gpu -q
grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -8 >> gpurun_out/round.log
run bench_gemm 300 python tools/bench_gemm.py
cat gpurun_out/bench_gemm.log >> gpurun_out/round.log
run bench_main 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
tail -1 gpurun_out/bench_main.log | cut -c1-1800 >> gpurun_out/round.log
cat gpurun_out/round.log
