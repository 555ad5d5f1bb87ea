#!/bin/bash
# compute-sanitizer memcheck over the GPU kernel tests (the heaviest GEMM shapes left out: cuBLAS references under the
# sanitizer take minutes) and, time permitting, the model tests.  Logs under gpurun_out/memcheck_*.log.
K='not (4096-4096-1024 or 2048-4096-512 or 1024-768-4096 or 2176-4096-1024 or 640-768-4096 or 800- or 896- or graph_replay or fresh_process)'
timeout ${1:-300} compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_kernels.log \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=line -k "$K" > gpurun_out/memcheck_kernels_pytest.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/memcheck_kernels_pytest.log; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_kernels.log
timeout ${2:-240} compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_model.log \
  python -m pytest tests/test_model_gpu.py -m gpu -q --tb=line -k "bf16 or training or lora or gpt4ts or edge" > gpurun_out/memcheck_model_pytest.log 2>&1
echo "model rc=$?"; tail -3 gpurun_out/memcheck_model_pytest.log; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_model.log
for f in gpurun_out/memcheck_kernels.log gpurun_out/memcheck_model.log; do head -c 200000 $f > $f.tmp; mv $f.tmp $f; done
