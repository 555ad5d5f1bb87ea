#!/bin/bash
# Bench line of every BASELINE config shape (configs[0] GPT4TS, [2] LUDB, [3] PSM, [4] Ventilator + LoRA); configs[1] is the headline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/configs.log
for wl in etth1_gpt4ts ludb_llama2_7b psm_gpt2_medium ventilator_llama2_7b; do
  echo "=== $wl" >> gpurun_out/configs.log
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $wl > gpurun_out/bench_$wl.log 2>&1; echo "exit $?" >> gpurun_out/configs.log
  grep '^{"metric"' gpurun_out/bench_$wl.log | cut -c1-5000 >> gpurun_out/configs.log
  grep -E "Error|error" gpurun_out/bench_$wl.log | tail -3 >> gpurun_out/configs.log
done
cat gpurun_out/configs.log
