#!/bin/bash
# Round-2 profile captures (ncu): launch lists of one forward / one training step, full captures of the layer-0 GEMMs, the
# attention forward and fused backward, and one TF32 GEMM of the evaluation parity mode.  Run through gpurun; then
#   python tools/summarize_ncu.py gpurun_out/r02_launches.csv r02_fwd
#   python tools/summarize_ncu.py gpurun_out/r02_launches_train.csv r02_train
#   python tools/summarize_ncu_full.py gpurun_out/r02_prof_gemm.ncu-rep r02_gemm_ncu "<title>"   (etc.)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-train --no-ref-gpu --no-extra-workloads"
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv"
timeout 900 ncu $M --log-file gpurun_out/r02_launches.csv $B --profile-step > gpurun_out/r02_ncu_fwd.log 2>&1
timeout 1200 ncu $M --log-file gpurun_out/r02_launches_train.csv $B --profile-step train > gpurun_out/r02_ncu_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 --profile-from-start off -s 4 -c 4 -f -o gpurun_out/r02_prof_gemm $B --profile-step > gpurun_out/r02_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ --profile-from-start off -s 1 -c 1 -f -o gpurun_out/r02_prof_attn $B --profile-step > gpurun_out/r02_ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd --profile-from-start off -s 1 -c 1 -f -o gpurun_out/r02_prof_attnbwd $B --profile-step train > gpurun_out/r02_ncu_attnbwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 2 -f -o gpurun_out/r02_prof_gemm_tf32 python tools/one_tf32_gemm.py > gpurun_out/r02_ncu_tf32.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches*.csv
