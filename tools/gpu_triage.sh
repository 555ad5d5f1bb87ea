#!/bin/bash
# Runs one GPU test file with -x, and re-runs it with every failing test deselected (a faulting kernel poisons the
# CUDA context of the process: the failures behind it say nothing). Logs under gpurun_out/triage_<name>_<i>.log.
# usage: tools/gpu_triage.sh <test file> <name> [max rounds]
f=$1; name=$2; max=${3:-4}
desel=()
for i in $(seq 1 $max); do
  log=gpurun_out/triage_${name}_$i.log
  timeout 900 python -m pytest "$f" -m gpu -x -q --tb=short "${desel[@]}" > $log 2>&1
  rc=$?
  tail -3 $log
  bad=$(grep -m1 '^FAILED\|^ERROR' $log | awk '{print $2}')
  if [ $rc -eq 0 ] || [ -z "$bad" ]; then break; fi
  echo "round $i: first failure $bad"
  desel+=(--deselect "$bad")
done
