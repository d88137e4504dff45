#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_cross_entropy" --timeout 240 --no-header -p no:cacheprovider > gpurun_out/m_ce.log 2>&1
echo "ce -> $?" | tee gpurun_out/model_summary.txt
for grp in test_fp32_logits test_bf16_logits test_gradients test_fp32_generation "test_bf16_generation or test_state_dict or test_align" test_midsize; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -k "$grp" --timeout 600 --no-header -p no:cacheprovider > "gpurun_out/m_${name}.log" 2>&1
  echo "$grp -> exit $?" | tee -a gpurun_out/model_summary.txt
  tail -5 "gpurun_out/m_${name}.log"
done
