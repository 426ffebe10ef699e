#!/bin/bash
# compute-sanitizer pass over the single-kernel C-ABI tests (small shapes).  A synccheck / memcheck error kills the CUDA
# context, so every test group runs in its own process.
# Usage (GPU box): bash tools/sanitize.sh [tools...] ; logs -> gpurun_out/sanitizer_<tool>_<group>.log, summary -> gpurun_out/sanitizer_summary.txt
mkdir -p gpurun_out
TOOLS=${@:-memcheck synccheck racecheck}
declare -A SEL
SEL[gemm]='test_gemm_bias_relu_ragged_rows or test_gemm_k384_layernorm or test_gemm_conv_taps_l2 or test_gemm_convert_broadcast or test_gemm_pair_kernel_forced or test_gemm_swish_and_glu or test_gemm_residual_dual_layernorm'
SEL[attn]='test_causal_attention or test_speaker_attention or test_spk_qkv_attn_fused'
SEL[ffn]='test_fused_ffn'
SEL[misc]='test_head or test_prep_input or test_embloss_kernel or test_dwconv_bn_swish_batch_and_one_step'
SEL[ret]='test_retention_chunkwise or test_retention_step'
SEL[p32]='test_p32_split_gemm or test_p32_linear_handle or test_p32_retention'
: > gpurun_out/sanitizer_summary.txt
for tool in $TOOLS; do
  for g in gemm attn ffn misc ret p32; do
    log=gpurun_out/sanitizer_${tool}_${g}.log
    timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool --print-limit 5 \
      python -m pytest tests/test_kernels_gpu.py tests/test_ls_gpu.py -q -k "${SEL[$g]}" --timeout 200 -p no:cacheprovider > $log 2>&1
    echo "== $tool / $g : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | pytest: $(grep -E ' passed| failed' $log | tail -1)" >> gpurun_out/sanitizer_summary.txt
  done
done
cat gpurun_out/sanitizer_summary.txt
