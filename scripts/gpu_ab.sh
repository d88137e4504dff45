#!/bin/bash
# A/B of two library builds on one box: GEMM kernel tests on the new build, then the C2 GEMM shapes + the train step
# with the new build and with multimodalanalytical_b200/lib/libmma_prev.so (MMA_B200_LIB override)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm or wgrad or grouped" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/k_gemm_ab.log 2>&1
echo "gemm tests -> $?"; tail -4 gpurun_out/k_gemm_ab.log
timeout 300 python scripts/gemm_diag.py > gpurun_out/gemm_diag_new.txt 2>&1; cat gpurun_out/gemm_diag_new.txt
PREV=multimodalanalytical_b200/lib/libmma_prev.so
if [ -f $PREV ]; then
  MMA_B200_LIB=$PREV timeout 300 python scripts/gemm_diag.py > gpurun_out/gemm_diag_prev.txt 2>&1; cat gpurun_out/gemm_diag_prev.txt
fi
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-decode --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new ', d['value'], d['ms_per_step'])"
  if [ -f $PREV ]; then
    MMA_B200_LIB=$PREV timeout 300 python bench.py --steps 30 --warmup 5 --no-decode --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev', d['value'], d['ms_per_step'])"
  fi
done
