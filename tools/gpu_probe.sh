#!/bin/bash
# Runs the GPU kernel tests group by group, each in its own process (a device-side trap poisons
# the CUDA context of the process that hit it), logging everything under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/nproc.txt
for grp in "gemm_store_bf16" "gemm_store_f32 or gemm_resid or gemm_gelu or gemm_swiglu or gemm_batched or gemm_rejects" "patch or revin" "rmsnorm or softmax or casts" "attn"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-40)
  echo "=== $grp" | tee -a gpurun_out/probe.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$grp" > "gpurun_out/probe_$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/probe.log
  tail -n 30 "gpurun_out/probe_$name.log" | tee -a gpurun_out/probe.log
done
