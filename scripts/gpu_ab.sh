#!/bin/bash
# A/B of two library builds on one box: kernel tests on the new build ($1 = pytest -k expression), then the train step
# with the new build and with multimodalanalytical_b200/lib/libmma_prev.so (MMA_B200_LIB override)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -k "${1:-gemm}" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/k_ab.log 2>&1
echo "tests -> $?"; tail -8 gpurun_out/k_ab.log
PREV=multimodalanalytical_b200/lib/libmma_prev.so
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-decode --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new ', d['value'], d['ms_per_step'])"
  if [ -f $PREV ]; then
    MMA_B200_LIB=$PREV timeout 300 python bench.py --steps 30 --warmup 5 --no-decode --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev', d['value'], d['ms_per_step'])"
  fi
done
