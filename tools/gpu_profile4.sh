#!/bin/bash
# ncu --set full captures of the decoder FFN, decoder time-attention and persistent GEMM kernels (one launch each)
mkdir -p gpurun_out
for spec in "ffn_kernel 4 prof_ffn_v4" "attn_kernel 4 prof_attn_v4" "gemm_persist_kernel 0 prof_gemmp_v4"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f \
      python tools/run_forward.py 1 > gpurun_out/ncu_$3.log 2>&1
done
ls -la gpurun_out | tail -8
