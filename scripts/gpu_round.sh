#!/bin/bash
# kernel tests + model tests + bench (+ optional ncu) in one GPU call
mkdir -p gpurun_out
rm -f gpurun_out/round_summary.txt
for grp in "test_gemm or test_grouped" "test_layernorm or test_embed or test_colsum" "test_attention or test_tc_attention or test_bf16_dh64" "test_cross_entropy or test_adam"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" --timeout 300 --no-header -p no:cacheprovider > "gpurun_out/k_${name}.log" 2>&1
  echo "kernels[$grp] -> exit $?" | tee -a gpurun_out/round_summary.txt
  grep -E "^(FAILED|ERROR)" "gpurun_out/k_${name}.log" | head -10 | tee -a gpurun_out/round_summary.txt
done
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/m_all.log 2>&1
echo "model -> exit $?" | tee -a gpurun_out/round_summary.txt
grep -E "^(FAILED|ERROR)" gpurun_out/m_all.log | head -20 | tee -a gpurun_out/round_summary.txt
tail -3 gpurun_out/m_all.log
timeout 900 python bench.py --steps 10 --warmup 3 $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench -> exit $?" | tee -a gpurun_out/round_summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$DO_NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python scripts/profile_step.py train > gpurun_out/prof_train.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py decode 64 > gpurun_out/prof_decode.log 2>&1
fi
