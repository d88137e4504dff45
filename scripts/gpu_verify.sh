#!/bin/bash
# final verification: every GPU test, smoke(), the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "gpu tests -> $?"; tail -3 gpurun_out/all_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench -> $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
