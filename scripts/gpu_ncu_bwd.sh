#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 400 $NCU -k "regex:^bwd_kernel$" -s 1 -c 3 -o gpurun_out/ncu_at5_bwd -f python scripts/profile_step.py train > gpurun_out/ncu5.log 2>&1
timeout 400 $NCU -k "regex:ln_bwd_kernel" -s 3 -c 2 -o gpurun_out/ncu_ln_bwd -f python scripts/profile_step.py train > gpurun_out/ncu6.log 2>&1
ls -la gpurun_out/ncu_at5_bwd.ncu-rep gpurun_out/ncu_ln_bwd.ncu-rep; tail -2 gpurun_out/ncu5.log
