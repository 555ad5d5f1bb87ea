#!/bin/bash
# Round-2 closing run: the whole GPU suite in ONE process (as the driver runs it), smoke, the default bench line, and the
# ncu launch lists / one full capture of the build with two epilogue warp sets + cluster split-K.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r02e_gpu_suite.log 2>&1; tail -3 gpurun_out/r02e_gpu_suite.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; tail -c 600 gpurun_out/r02e_bench_n1.json
B="python bench.py --no-cpu-baseline --no-train --no-ref-gpu --no-extra-workloads --no-precision-modes"
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv"
timeout 300 ncu $M --log-file gpurun_out/r02e_launches.csv $B --profile-step > gpurun_out/r02e_ncu_fwd.log 2>&1
timeout 300 ncu $M --log-file gpurun_out/r02e_launches_psm.csv $B --workload psm_gpt2_medium --profile-step > gpurun_out/r02e_ncu_psm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_nt_kernel --profile-from-start off -s 6 -c 3 -f -o gpurun_out/r02e_prof_gemm_psm $B --workload psm_gpt2_medium --profile-step > gpurun_out/r02e_ncu_gemm_psm.log 2>&1
ls -la gpurun_out/r02e_*
