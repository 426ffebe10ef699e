#!/bin/bash
# Round-end artefacts: full GPU suite, bench line, secondary measurements, ncu launch list of one forward and
# ncu --set full of the top kernels.  usage: tools/gpu_final.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json | head -8
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
FSEEND_EXTRA_LONG=1 timeout 500 python tools/bench_extra.py > gpurun_out/${TAG}_extra.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_extra.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 33 -c 33 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/run_forward.py 2 > gpurun_out/${TAG}_ncu_launch.log 2>&1
for spec in "ffn_pair_kernel 4 ffn" "attn2_kernel 4 attn" "spkfuse_kernel 0 spkfuse" "gemm_pair_kernel 6 gemmpair"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/${TAG}_$3 -f \
      python tools/run_forward.py 1 > gpurun_out/${TAG}_ncu_$3.log 2>&1
done
ls gpurun_out | grep ${TAG}
