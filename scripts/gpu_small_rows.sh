#!/bin/bash
# row-blocked small decode path: kernel test, parity test, then ms/step with the path limited to 64 rows and extended to 512
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_r2_gpu.py -m gpu -q -k "small_linear or row_blocked" --timeout 200 --no-header -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -15
S="8x10 10x10 16x10 24x10 32x10 48x10 3x30 8x30 16x30"
MMA_DECODE_SMALL_ROWS=64 timeout 200 python scripts/decode_bench.py $S --gated 2>/dev/null | tail -1
MMA_DECODE_SMALL_ROWS=512 timeout 200 python scripts/decode_bench.py $S --gated 2>/dev/null | tail -1
