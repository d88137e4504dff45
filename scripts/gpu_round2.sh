#!/bin/bash
# Round-2 evidence run on one B200: GPU test suite, smoke, bench (ours + reference arm), ncu launch lists and
# `--set full` captures of the kernels added / changed this round.  Outputs under gpurun_out/ (summarised into profiles/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err
L="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
$L --log-file gpurun_out/r2_launches_train.csv python scripts/profile_step.py train > gpurun_out/r2_p1.log 2>&1
$L --log-file gpurun_out/r2_launches_train_c2paper.csv python scripts/profile_case.py c2_paper > gpurun_out/r2_p2.log 2>&1
$L --log-file gpurun_out/r2_launches_train_c4.csv python scripts/profile_case.py c4 > gpurun_out/r2_p3.log 2>&1
$L -s 4560 -c 456 --log-file gpurun_out/r2_launches_decode.csv python scripts/profile_step.py decode 256 128 > gpurun_out/r2_p4.log 2>&1
F="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 300 $F -k regex:gemm2_kernel -s 20 -c 14 -o gpurun_out/r2_ncu_gemm2 -f python scripts/profile_step.py train > gpurun_out/r2_n1.log 2>&1
if [ "${FULL_NCU:-0}" = "1" ]; then
timeout 300 $F -k "regex:glu_fwd_kernel|dglu_kernel" -s 2 -c 4 -o gpurun_out/r2_ncu_glu -f python scripts/profile_case.py c2_paper > gpurun_out/r2_n2.log 2>&1
timeout 300 $F -k "regex:fwd_blk_kernel|bwd_blk_kernel|merge_fwd|merge_bwd" -s 0 -c 4 -o gpurun_out/r2_ncu_attnblk -f python scripts/profile_case.py c4 > gpurun_out/r2_n3.log 2>&1
fi
timeout 300 $F -k "regex:decode_self_attn2|beam_step|gemm_tc_kernel" -s 1098 -c 6 -o gpurun_out/r2_ncu_decode -f python scripts/profile_step.py decode 256 128 > gpurun_out/r2_n4.log 2>&1
# the one-launch decode step (4 spectra x 10 beams): launch list of a few steps, one full capture, per-phase stamps
$L -c 60 --log-file gpurun_out/r2_launches_decode_b4.csv python scripts/profile_step.py decode 4 24 > gpurun_out/r2_p5.log 2>&1
timeout 300 $F -k regex:decode_step_kernel -s 10 -c 1 -o gpurun_out/r2_ncu_decode_step -f python scripts/profile_step.py decode 4 24 > gpurun_out/r2_n5.log 2>&1
python scripts/decode_step_phases.py 1x10 4x10 --gated > gpurun_out/r2_decode_step_phases.txt 2>&1
ls -la gpurun_out/r2_ncu_*.ncu-rep
