#!/bin/bash
# Quick synccheck + racecheck pass (memcheck: tools/sanitize.sh, clean in r02) with hard per-process limits.
mkdir -p gpurun_out
declare -A SEL
SEL[gemm]='test_gemm_bias_relu_ragged_rows or test_gemm_k384_layernorm or test_gemm_conv_taps_l2 or test_gemm_convert_broadcast or test_gemm_pair_kernel_forced or test_gemm_swish_and_glu or test_gemm_residual_dual_layernorm'
SEL[attn]='(test_causal_attention and (130 or 257 or 64 or 129)) or test_speaker_attention or test_spk_qkv_attn_fused'
SEL[ffn]='test_fused_ffn'
SEL[misc]='test_head or test_prep_input or test_embloss_kernel or test_dwconv_bn_swish_batch_and_one_step or test_retention_chunkwise or test_retention_step'
SEL[p32]='test_p32_split_gemm or test_p32_linear_handle or (test_p32_retention and 384)'
: > gpurun_out/sanitizer_summary.txt
for tool in synccheck racecheck; do
  for g in gemm attn ffn misc p32; do
    log=gpurun_out/sanitizer_${tool}_${g}.log
    timeout ${SAN_TIMEOUT:-150} compute-sanitizer --tool $tool --print-limit 3 \
      python -m pytest tests/test_kernels_gpu.py tests/test_ls_gpu.py -q -k "${SEL[$g]}" --timeout 140 -p no:cacheprovider > $log 2>&1
    echo "== $tool / $g : rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | pytest: $(grep -E ' passed| failed' $log | tail -1)" >> gpurun_out/sanitizer_summary.txt
  done
done
cat gpurun_out/sanitizer_summary.txt
