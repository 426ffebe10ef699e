#!/bin/bash
# Round profile: bench line, ncu launch list of the same forward, ncu --set full of the top kernels.
mkdir -p gpurun_out
python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 6000 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 47 -c 47 --csv --log-file gpurun_out/launches.csv \
    python tools/run_forward.py 2 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 23 -c 2 -o gpurun_out/prof_gemm_ffn \
    python tools/run_forward.py 1 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 4 -c 1 -o gpurun_out/prof_attn \
    python tools/run_forward.py 1 > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out
