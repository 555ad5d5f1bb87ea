#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name" >> gpurun_out/round.log
  timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log
}
run test_kernels_fwd 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not bwd and not scalar_epilogue and not colsum"
tail -3 gpurun_out/test_kernels_fwd.log >> gpurun_out/round.log
run test_kernels_bwd 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "bwd or scalar_epilogue or colsum"
grep -E "passed|failed|Error|assert" gpurun_out/test_kernels_bwd.log | tail -15 >> gpurun_out/round.log
run test_model_fwd 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "not training"
grep -E "parity|passed|failed|Error" gpurun_out/test_model_fwd.log | tail -12 >> gpurun_out/round.log
run test_model_train 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "training"
grep -E "parity|passed|failed|Error|error" gpurun_out/test_model_train.log | tail -20 >> gpurun_out/round.log
run bench_gemm 300 python tools/bench_gemm.py
cat gpurun_out/bench_gemm.log >> gpurun_out/round.log
run bench_psm 600 python bench.py --workload psm_gpt2_medium --steps 20 --warmup 3
tail -1 gpurun_out/bench_psm.log >> gpurun_out/round.log
run bench_main 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
tail -1 gpurun_out/bench_main.log >> gpurun_out/round.log
cat gpurun_out/round.log
