#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the kernels that carry hand-rolled mbarrier / TMEM / TMA hand-offs:
# CTA-pair GEMM (all epilogues), fused gate kernels, grouped wgrad, tcgen05 attention fwd/bwd, pipelined LayerNorm backward,
# decode self-attention.  One parametrisation each (the sanitizer slows kernels ~50x).
mkdir -p gpurun_out
SEL='(test_gemm_pair_kernel and 12416) or (test_ffn_glu_pair_kernels and 12300) or test_grouped_wgrad_with_fused_bias_grad or (test_bf16_dh64_attention_both_tensor_core_paths and tcgen05 and 36-36) or (test_layernorm_bwd_prefetching_variant) or test_c5_beam10'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py tests/test_configs_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2_sanitizer_summary.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$tool.log | tail -3 | tee -a gpurun_out/r2_sanitizer_summary.txt
done
