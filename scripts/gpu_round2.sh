#!/bin/bash
# full GPU round: all gpu tests, bench, decode sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "gpu tests -> $?"; tail -6 gpurun_out/all_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench -> $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$DO_SWEEP" ]; then timeout 600 python scripts/decode_sweep.py ${SWEEP:-1,8,64,256,1024} 10 2>&1 | tee gpurun_out/decode_sweep.txt; fi
