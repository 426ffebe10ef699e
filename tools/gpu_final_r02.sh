#!/bin/bash
# Round-2 artefacts (every step with a hard timeout): full GPU suite, bench line (driver arguments), reference arm,
# ncu launch list of one FS forward, ncu --set full of the decoder attention launch and of the persistent parity GEMM.
TAG=${1:-r02_final}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 200 --warmup 10 --no-secondary --no-cpu-baseline > gpurun_out/${TAG}_bench_200.json 2>> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 35 -c 35 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/fs_forward_once.py 2 > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn3 -s 2 -c 1 -o gpurun_out/${TAG}_attn3 -f \
    python tools/fs_forward_once.py 2 > gpurun_out/${TAG}_ncu_attn3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:p32_gemm_persist -s 86 -c 1 -o gpurun_out/${TAG}_p32persist -f \
    python tools/ls_forward_once.py fp32 2 > gpurun_out/${TAG}_ncu_p32.log 2>&1
ls -la gpurun_out | grep ${TAG}
