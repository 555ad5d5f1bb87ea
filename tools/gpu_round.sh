#!/bin/bash
# Full confirmation round: GPU tests, bench lines (BIDMC headline + PSM GPT-2), ncu launch lists (fwd + train), full captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
grep -E "parity\]|passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -40 >> gpurun_out/round.log
run bench_main 1200 python bench.py --steps 10 --warmup 3
tail -1 gpurun_out/bench_main.log | cut -c1-6000 >> gpurun_out/round.log
run bench_psm 600 python bench.py --steps 20 --warmup 3 --workload psm_gpt2_medium
tail -1 gpurun_out/bench_psm.log | cut -c1-6000 >> gpurun_out/round.log
run ncu_fwd 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu-baseline --no-train --no-ref-gpu
run ncu_train 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python bench.py --profile-step train --no-cpu-baseline --no-train --no-ref-gpu
run ncu_full_gemm 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 --profile-from-start off -s 4 -c 4 -f -o gpurun_out/prof_gemm python bench.py --profile-step --no-cpu-baseline --no-train --no-ref-gpu
run ncu_full_attn 900 ncu --set full --clock-control none --import-source on -k regex:attn_ --profile-from-start off -s 1 -c 1 -f -o gpurun_out/prof_attn python bench.py --profile-step --no-cpu-baseline --no-train --no-ref-gpu
cat gpurun_out/round.log
