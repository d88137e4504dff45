#!/bin/bash
# end-of-round evidence: all GPU tests, smoke, bench (+reference arm), launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "gpu tests -> $?"; tail -3 gpurun_out/all_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench -> $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference -> $?"; cat gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python scripts/profile_step.py train > gpurun_out/prof_train.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py decode 256 > gpurun_out/prof_decode.log 2>&1
echo done
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ln_bwd -s 4 -c 3 -o gpurun_out/ncu_ln_bwd_v6 -f python scripts/profile_step.py train > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/ncu_ln_bwd_v6.ncu-rep
