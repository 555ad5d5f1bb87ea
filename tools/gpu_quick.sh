#!/bin/bash
# Quick round: GPU tests + headline bench line (+ optional train launch list with "train").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/round.log
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> gpurun_out/round.log; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/round.log; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -20 >> gpurun_out/round.log
run bench_main 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu
tail -1 gpurun_out/bench_main.log | cut -c1-3500 >> gpurun_out/round.log
if [ "$1" = "train" ]; then
run ncu_train 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python bench.py --profile-step train --no-cpu-baseline --no-train --no-ref-gpu
fi

if [ "$2" = "attnbwd" ]; then
run ncu_full_attnbwd 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd --profile-from-start off -s 0 -c 3 -f -o gpurun_out/prof_attnbwd python bench.py --profile-step train --no-cpu-baseline --no-train --no-ref-gpu
echo "attnbwd capture done" >> gpurun_out/round.log
fi
cat gpurun_out/round.log
