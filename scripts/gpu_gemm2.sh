#!/bin/bash
mkdir -p gpurun_out
MMA_GEMM2_VERBOSE=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "test_gemm_pair_kernel" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/gemm2_test_swz1.log 2>&1
echo "swz1 -> $?"; tail -25 gpurun_out/gemm2_test_swz1.log
MMA_GEMM2_SWZ=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "test_gemm_pair_kernel" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/gemm2_test_swz0.log 2>&1
echo "swz0 -> $?"; tail -8 gpurun_out/gemm2_test_swz0.log
for d in 0 1 2 3; do
  MMA_GEMM_DBG=$d timeout 300 python scripts/gemm_diag.py > gpurun_out/gemm2_diag_$d.txt 2>&1
  cat gpurun_out/gemm2_diag_$d.txt
done
MMA_GEMM2=0 timeout 300 python scripts/gemm_diag.py > gpurun_out/gemm2_diag_old.txt 2>&1
cat gpurun_out/gemm2_diag_old.txt
