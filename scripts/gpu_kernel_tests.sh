#!/bin/bash
# Runs the kernel unit tests in separate processes (a trapped kernel poisons only its own group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in test_gemm_fwd_kmajor test_gemm_dgrad test_gemm_wgrad test_gemm_epilogues test_gemm_dropout \
           "test_layernorm or test_embed or test_colsum" test_attention "test_cross_entropy or test_adam"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" --timeout 240 -x --no-header -p no:cacheprovider \
      > "gpurun_out/k_${name}.log" 2>&1
  echo "$grp -> exit $?" | tee -a gpurun_out/kernel_summary.txt
  tail -3 "gpurun_out/k_${name}.log"
done
