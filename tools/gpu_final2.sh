#!/bin/bash
# Last confirmation of the round: the whole GPU suite in one process, smoke, the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r02f_gpu_suite.log 2>&1; tail -3 gpurun_out/r02f_gpu_suite.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; tail -c 300 gpurun_out/r02f_bench_n1.json
