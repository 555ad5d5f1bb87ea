#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run test_2cta 120 python tools/test_2cta.py
tail -12 gpurun_out/test_2cta.log >> gpurun_out/round.log
run test_gemm_1cta 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k gemm
tail -2 gpurun_out/test_gemm_1cta.log >> gpurun_out/round.log
if grep -q 2CTA_OK gpurun_out/test_2cta.log; then
  MTS_GEMM_2CTA=1 timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm_2cta.log 2>&1
  echo "=== bench_gemm 2cta" >> gpurun_out/round.log; cat gpurun_out/bench_gemm_2cta.log >> gpurun_out/round.log
  run bench_main_2cta 900 env MTS_GEMM_2CTA=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train
  tail -1 gpurun_out/bench_main_2cta.log | cut -c1-300 >> gpurun_out/round.log
fi
run bench_gemm_1cta 300 python tools/bench_gemm.py
cat gpurun_out/bench_gemm_1cta.log >> gpurun_out/round.log
run bench_main_1cta 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train
tail -1 gpurun_out/bench_main_1cta.log | cut -c1-300 >> gpurun_out/round.log
cat gpurun_out/round.log
