#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:at5 -c 4 -o gpurun_out/ncu_at5 -f python scripts/profile_step.py train > gpurun_out/ncu1.log 2>&1
$NCU -k regex:gemm_tc_kernel.*1.*1.*6 -s 4 -c 4 -o gpurun_out/ncu_wgrad -f python scripts/profile_step.py train > gpurun_out/ncu2.log 2>&1
$NCU -k regex:ln_bwd_kernel -s 3 -c 2 -o gpurun_out/ncu_ln_bwd -f python scripts/profile_step.py train > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
