#!/bin/bash
# one `ncu --set full` capture per hot kernel of the C2 training step (1 GPU; never wrap a multi-rank command)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 500 $NCU -k regex:gemm2_kernel -s 20 -c 14 -o gpurun_out/ncu_gemm2 -f python scripts/profile_step.py train > gpurun_out/ncu1.log 2>&1
timeout 300 $NCU -k regex:wgrad2_group_kernel -s 1 -c 2 -o gpurun_out/ncu_wgrad2 -f python scripts/profile_step.py train > gpurun_out/ncu2.log 2>&1
timeout 300 $NCU -k "regex:^(fwd_kernel|bwd_pipe_kernel)$" -s 2 -c 4 -o gpurun_out/ncu_at5 -f python scripts/profile_step.py train > gpurun_out/ncu3.log 2>&1
timeout 300 $NCU -k "regex:ln_bwd|ln_fwd_kernel|adam_kernel" -s 8 -c 6 -o gpurun_out/ncu_rowops -f python scripts/profile_step.py train > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
