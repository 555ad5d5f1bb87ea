#!/bin/bash
# GPU test suite only (optionally: -k expression), plus smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|^E  |grad parity\]|parity\]" gpurun_out/pytest_gpu.log | cut -c1-1500 | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
