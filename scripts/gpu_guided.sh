#!/bin/bash
# guided decoding / logits-processor GPU tests + the decode tests that share the step kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_guided.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/guided.log 2>&1
echo "guided -> $?"; tail -25 gpurun_out/guided.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_configs_gpu.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "generation or decode or greedy or beam" > gpurun_out/decode_regress.log 2>&1
echo "decode regress -> $?"; tail -5 gpurun_out/decode_regress.log
