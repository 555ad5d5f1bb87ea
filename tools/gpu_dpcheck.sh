#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check exit $?"
grep -E "dp_check\]|Error" gpurun_out/dp_check.log | tail -8
