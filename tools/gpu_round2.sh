#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest model" > gpurun_out/round.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s > gpurun_out/test_model.log 2>&1; echo "exit $?" >> gpurun_out/round.log
grep -E "parity|passed|failed|Error|error" gpurun_out/test_model.log | tail -20 >> gpurun_out/round.log
echo "=== gemm kernels tests" >> gpurun_out/round.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k gemm > gpurun_out/test_gemm.log 2>&1; echo "exit $?" >> gpurun_out/round.log
tail -3 gpurun_out/test_gemm.log >> gpurun_out/round.log
echo "=== bench_gemm" >> gpurun_out/round.log
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "exit $?" >> gpurun_out/round.log
cat gpurun_out/bench_gemm.log >> gpurun_out/round.log
echo "=== ncu launch list (one profiled step)" >> gpurun_out/round.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "exit $?" >> gpurun_out/round.log
echo "=== ncu full (layer-0 GEMMs)" >> gpurun_out/round.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 --profile-from-start off -s 4 -c 4 -f -o gpurun_out/prof_gemm python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "exit $?" >> gpurun_out/round.log
echo "=== bench psm" >> gpurun_out/round.log
timeout 600 python bench.py --workload psm_gpt2_medium --steps 20 --warmup 3 > gpurun_out/bench_psm.log 2>&1; echo "exit $?" >> gpurun_out/round.log
tail -1 gpurun_out/bench_psm.log >> gpurun_out/round.log
echo "=== bench headline" >> gpurun_out/round.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_main.log 2>&1; echo "exit $?" >> gpurun_out/round.log
tail -1 gpurun_out/bench_main.log >> gpurun_out/round.log
cat gpurun_out/round.log
