#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag1_gpu.txt
for d in 0 1 2 3 4 5; do
  MMA_GEMM_DBG=$d timeout 300 python scripts/gemm_diag.py > gpurun_out/gemm_diag_$d.txt 2>&1
done
timeout 900 python scripts/decode_sweep.py 1,8,64,256,1024 10 > gpurun_out/decode_sweep.txt 2>&1
cat gpurun_out/gemm_diag_*.txt gpurun_out/decode_sweep.txt
