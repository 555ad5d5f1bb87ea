#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_psm.csv python bench.py --profile-step --no-cpu-baseline --no-train --no-ref-gpu --workload psm_gpt2_medium > gpurun_out/ncu_psm.log 2>&1
echo "ncu exit $?"
for pdl in 1 0; do
MTS_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 3 --workload psm_gpt2_medium --no-cpu-baseline --no-ref-gpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$pdl psm fwd', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
MTS_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$pdl bidmc fwd', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
done
