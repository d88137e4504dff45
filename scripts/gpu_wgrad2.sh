#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "test_grouped or test_gemm_pair" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/wgrad2_test.log 2>&1
echo "tests -> $?"; tail -15 gpurun_out/wgrad2_test.log
MMA_WGRAD2=1 timeout 300 python scripts/wgrad_bench.py 2>&1 | tee gpurun_out/wgrad_bench_new.txt
MMA_WGRAD2=0 timeout 300 python scripts/wgrad_bench.py 2>&1 | tee gpurun_out/wgrad_bench_old.txt
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/m_all.log 2>&1
echo "model -> $?"; tail -5 gpurun_out/m_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-decode 2>&1 | tee gpurun_out/bench_wgrad2.json
MMA_WGRAD2=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-decode 2>&1 | tee gpurun_out/bench_wgrad_old.json
