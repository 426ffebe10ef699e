#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "ffn" 2>&1 | tail -3
python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err
ncu --set full --clock-control none --import-source on -k regex:ffn_kernel -s 4 -c 1 -o gpurun_out/prof_ffn \
    python tools/run_forward.py 1 > gpurun_out/ncu_ffn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 4 -c 2 -o gpurun_out/prof_attn2 \
    python tools/run_forward.py 1 > gpurun_out/ncu_attn2.log 2>&1
ls gpurun_out
