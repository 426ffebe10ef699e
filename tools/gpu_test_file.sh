#!/bin/bash
# run each test function of one file in its own process (isolates trapped kernels)
F=$1
TESTS=$(python -m pytest $F -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
for t in $TESTS; do
  echo "=== $t"
  timeout 600 python -m pytest "$t" -x -q -s 2>&1 | grep -vE "^$" | tail -${TAIL:-14}
done
