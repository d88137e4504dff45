#!/bin/bash
# A/B of the current build against multimodalanalytical_b200/lib/libmma_prev.so on one box: GPU tests on the new build,
# GEMM shapes, then the C2 / C2-paper / C4 train steps with both builds, interleaved
mkdir -p gpurun_out
PREV=$PWD/multimodalanalytical_b200/lib/libmma_prev.so
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --no-header -p no:cacheprovider > gpurun_out/ab2_tests.log 2>&1
echo "tests -> $?"; tail -4 gpurun_out/ab2_tests.log
echo "--- new"; python scripts/gemm_diag.py 2>&1 | grep -v max_ctas
echo "--- prev"; MMA_B200_LIB=$PREV python scripts/gemm_diag.py 2>&1 | grep -v max_ctas
for i in 1 2; do
  python bench.py --steps 30 --warmup 5 --no-decode --no-cpu --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new ', d['value'], d['ms_per_step'], d['roofline']['us_per_launch'], d['last_loss'])"
  MMA_B200_LIB=$PREV python bench.py --steps 30 --warmup 5 --no-decode --no-cpu --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev', d['value'], d['ms_per_step'], d['roofline']['us_per_launch'], d['last_loss'])"
done
python scripts/case_bench.py c2_paper c4
MMA_B200_LIB=$PREV python scripts/case_bench.py c2_paper c4
