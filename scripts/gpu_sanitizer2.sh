#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: the one-launch decode step (cluster barriers, distributed
# shared memory, cp.async staging), the small-row LayerNorm + product kernel, the 4-CTA-cluster residual + LayerNorm kernel.
mkdir -p gpurun_out
cat > /tmp/san_decode.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
from multimodalanalytical_b200.wrapper import HFWrapper
from multimodalanalytical_b200 import decode as dec
c = dict(bench.C2)
m = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=10, precision="bf16", seed=1,
              **bench.model_kwargs(c, **bench.PAPER))
m.eval()
m.generation_config["max_length"] = 6
batch = bench.map_batch(bench.synth_batch(c, 2, 5), lambda x: x.cuda())
out = m.generate(batch, n_beams=10, use_graph=False)        # one-launch step (2 clusters)
assert m.generator._persist_plan(m.generator._states[(2, 10, 6)]) is not None
dec.PERSIST_DECODE = False
out2 = m.generate(batch, n_beams=10, use_graph=False)       # small_linear path (20 rows)
torch.cuda.synchronize()
print("ok", tuple(out.shape), tuple(out2.shape))
PY
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 python /tmp/san_decode.py > gpurun_out/r2_sanitizer2_decode_$tool.log 2>&1
  echo "decode_step + small_linear: $tool rc=$?" | tee -a gpurun_out/r2_sanitizer2_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok" gpurun_out/r2_sanitizer2_decode_$tool.log | tail -3 | tee -a gpurun_out/r2_sanitizer2_summary.txt
done
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "(test_gemm_resid_layernorm_fused and 200-512) or (test_small_linear_decode_products and 30)" > gpurun_out/r2_sanitizer2_kern_$tool.log 2>&1
  echo "resid+LN cluster kernel, small_linear test: $tool rc=$?" | tee -a gpurun_out/r2_sanitizer2_summary.txt
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer2_kern_$tool.log | tail -3 | tee -a gpurun_out/r2_sanitizer2_summary.txt
done
