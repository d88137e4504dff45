#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k "regex:^(fwd_kernel|bwd_kernel)$" -c 4 -o gpurun_out/ncu_at5 -f python scripts/profile_step.py train > gpurun_out/ncu1.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 30 -c 12 -o gpurun_out/ncu_gemm_train -f python scripts/profile_step.py train > gpurun_out/ncu2.log 2>&1
$NCU -k regex:wgrad_group_kernel -s 2 -c 2 -o gpurun_out/ncu_wgrad -f python scripts/profile_step.py train > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
