#!/bin/bash
mkdir -p gpurun_out
# decode steps around t = 60..64 at a large batch (1024 spectra x 10 beams): per-kernel device time
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -s 4500 -c 300 --csv --log-file gpurun_out/launches_decode_b1024.csv python scripts/profile_step.py decode 1024 66 > gpurun_out/prof_decode_b1024.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -s 4500 -c 300 --csv --log-file gpurun_out/launches_decode_b64.csv python scripts/profile_step.py decode 64 66 > gpurun_out/prof_decode_b64.log 2>&1
tail -2 gpurun_out/prof_decode_b1024.log
