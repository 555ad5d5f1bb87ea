#!/bin/bash
# Shared-prefix round: GPU test suite, bench line (with per-sample-prompts + HF GPU legs), ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
tail -15 gpurun_out/pytest_gpu.log >> gpurun_out/round.log
run bench_main 1200 python bench.py --steps 10 --warmup 3
tail -1 gpurun_out/bench_main.log | cut -c1-6000 >> gpurun_out/round.log
run ncu_fwd 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu-baseline --no-train --no-ref-gpu
cat gpurun_out/round.log
