#!/bin/bash
# One GPU call: GPU test suite, bench line, ncu launch list of one forward, ncu --set full of the top kernels.
# usage: tools/gpu_round.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 35 -c 35 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/run_forward.py 2 > gpurun_out/${TAG}_ncu_launch.log 2>&1
for spec in "ffn_pair_kernel 4 ffn" "attn_kernel 4 attn" "gemm_pair_kernel 2 gemmpair"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/${TAG}_$3 -f \
      python tools/run_forward.py 1 > gpurun_out/${TAG}_ncu_$3.log 2>&1
done
ls -la gpurun_out | tail -12
