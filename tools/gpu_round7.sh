#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run test_gemm 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm"
grep -E "passed|failed|Error|assert" gpurun_out/test_gemm.log | tail -8 >> gpurun_out/round.log
run test_lora 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "lora"
grep -E "parity|passed|failed|Error|error" gpurun_out/test_lora.log | tail -12 >> gpurun_out/round.log
run test_model 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -k "not lora"
grep -E "passed|failed|Error" gpurun_out/test_model.log | tail -8 >> gpurun_out/round.log
cat gpurun_out/round.log
