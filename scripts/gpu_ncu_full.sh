#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:gemm_tc_kernel -s 60 -c 6 -o gpurun_out/ncu_gemm_train -f python scripts/profile_step.py train > gpurun_out/ncu1.log 2>&1
$NCU -k regex:fwd_kernel -c 2 -o gpurun_out/ncu_attn_fwd -f python scripts/profile_step.py train > gpurun_out/ncu2.log 2>&1
$NCU -k regex:bwd_dkv_kernel -c 1 -o gpurun_out/ncu_attn_dkv -f python scripts/profile_step.py train > gpurun_out/ncu3.log 2>&1
$NCU -k regex:ln_bwd_kernel -s 3 -c 2 -o gpurun_out/ncu_ln_bwd -f python scripts/profile_step.py train > gpurun_out/ncu4.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 40 -c 8 -o gpurun_out/ncu_gemm_decode -f python scripts/profile_step.py decode 64 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
