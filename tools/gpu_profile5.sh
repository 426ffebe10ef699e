#!/bin/bash
# ncu --set full capture: KERNEL (regex), SKIP launches, COUNT, OUT name
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 -f \
    python tools/run_forward.py 1 > gpurun_out/ncu_$4.log 2>&1
ls -la gpurun_out/$4.ncu-rep
