#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 500 $NCU -k "regex:decode_attn_kernel" -s 360 -c 2 -o gpurun_out/ncu_decode_attn -f python scripts/profile_step.py decode 1024 66 > gpurun_out/ncu7.log 2>&1
ls -la gpurun_out/ncu_decode_attn.ncu-rep; tail -2 gpurun_out/ncu7.log
