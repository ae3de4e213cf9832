#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel kills only that file's context),
# with a hard timeout, and collects the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc_all=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -q -m gpu -rA --tb=short -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  rc=$?
  echo "== $name exit $rc"
  tail -n 40 "gpurun_out/${name}.log" | grep -E "passed|failed|error|PASSED|FAILED|ERROR|rel err|timeout|Assert" | tail -n 30
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
