#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 600 python tools/bench_gemm.py --shared-prefix-only 2>&1 | grep "^{" | cut -c1-220
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bidmc fwd', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d['train_step']['value'], d['train_step']['ms_per_step'], d['clocks'], 'roof', d['roofline']['achieved'])"
