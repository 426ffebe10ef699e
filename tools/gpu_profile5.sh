#!/bin/bash
# ncu --set full capture of one kernel: KERNEL (regex), SKIP launches, OUT name
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f \
    python tools/run_forward.py 1 > gpurun_out/ncu_$3.log 2>&1
ls -la gpurun_out/$3.ncu-rep
