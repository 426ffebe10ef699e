#!/bin/bash
# Quick GPU call: selected tests (pytest -k "$2", default all GPU tests) + one bench line.  usage: tools/gpu_quick.sh TAG [KEXPR]
TAG=${1:-q}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q -s -k "$2" ) > gpurun_out/${TAG}_pytest.log 2>&1
else
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
fi
tail -15 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json
