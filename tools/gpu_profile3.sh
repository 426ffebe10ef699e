#!/bin/bash
# Final-state profile of the bench forward: launch list (shares) + ncu --set full of the dominant kernel (fused FFN, TS).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 35 -c 35 --csv --log-file gpurun_out/launches_v3.csv \
    python tools/run_forward.py 2 > gpurun_out/ncu_launch3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffn_kernel -s 4 -c 1 -o gpurun_out/prof_ffn_ts \
    python tools/run_forward.py 1 > gpurun_out/ncu_ffn_ts.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 16 -c 4 -o gpurun_out/prof_gemm_dec \
    python tools/run_forward.py 1 > gpurun_out/ncu_gemm_dec.log 2>&1
ls gpurun_out
