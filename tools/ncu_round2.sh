#!/bin/bash
# ncu --set full captures the round-1 review asked for (one GPU, short commands): the decoder-sized gemm_pair launch,
# the LS kernels (retention, chunk state, depthwise conv) and the parity-path GEMM.  Reports -> gpurun_out/*.ncu-rep
mkdir -p gpurun_out
# FS forward, launch ids as in profiles/r01_final_launches.csv: warm-up forward = 33 launches, decoder out1_ln = id 23
ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 17 -c 4 -o gpurun_out/r02_gemmpair_dec \
    python tools/fs_forward_once.py 2 > gpurun_out/ncu_fs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"retention_kernel|ret_chunk_state|dwconv16" -s 16 -c 4 \
    -o gpurun_out/r02_ls_fp16_kernels python tools/ls_forward_once.py fp16 2 > gpurun_out/ncu_ls.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"p32_retention|p32_spk" -s 6 -c 3 \
    -o gpurun_out/r02_ls_p32_kernels python tools/ls_forward_once.py fp32 2 >> gpurun_out/ncu_ls.log 2>&1
ls -la gpurun_out/*.ncu-rep
