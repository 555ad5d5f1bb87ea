#!/bin/bash
# 2-GPU data-parallel check: gradient equivalence over NCCL + the bench line at N=2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/dp.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check exit $?" >> gpurun_out/dp.log
grep dp_check gpurun_out/dp_check.log >> gpurun_out/dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_dp$N.log 2>&1; echo "bench exit $?" >> gpurun_out/dp.log
grep '^{"metric"' gpurun_out/bench_dp$N.log | cut -c1-4000 >> gpurun_out/dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --workload psm_gpt2_medium > gpurun_out/bench_psm_dp$N.log 2>&1; echo "bench psm exit $?" >> gpurun_out/dp.log
grep '^{"metric"' gpurun_out/bench_psm_dp$N.log | cut -c1-4000 >> gpurun_out/dp.log
cat gpurun_out/dp.log
