#!/bin/bash
# Run every GPU test function in its own process (a trapped kernel poisons the CUDA context of its process only),
# with a timeout each; results go to gpurun_out/diag.log.
mkdir -p gpurun_out
LOG=gpurun_out/diag.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv >> $LOG 2>&1
python __graft_entry__.py >> $LOG 2>&1
TESTS=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
for t in $TESTS; do
  echo "=== $t" >> $LOG
  timeout 600 python -m pytest "$t" -x -q -s 2>&1 | grep -v "^$" | tail -${TAIL:-25} >> $LOG
  echo "=== exit $?" >> $LOG
done
grep -E "^===|passed|failed|error" $LOG | tail -80
