#!/bin/bash
# compute-sanitizer memcheck over what the last session of round 2 added: small_linear in row blocks (rows > 64), the
# derivative patch kernel (device preprocessor + row-indexed collate), the LN(x + f(x)) layer order (train step + decode)
mkdir -p gpurun_out
cat > /tmp/san_post.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from tests.helpers import load_case
from tests.test_model_gpu import build
fx = load_case("post_ln")
m = build(fx, "bf16", dropout=0.1)
m.train()
out = m.forward(fx["batch"]); out.loss.backward()
m.eval()
seq = m.generate(fx["batch"], n_beams=3, use_graph=False)
torch.cuda.synchronize()
print("ok", float(out.loss), tuple(seq.shape))
PY
S=gpurun_out/r2_sanitizer3_summary.txt; : > $S
timeout 150 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 python /tmp/san_post.py > gpurun_out/r2_sanitizer3_postln.log 2>&1
echo "post-LN train step + beam-3 decode: memcheck rc=$?" | tee -a $S
grep -E "ERROR SUMMARY|^ok" gpurun_out/r2_sanitizer3_postln.log | tail -2 | tee -a $S
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py tests/test_patches.py tests/test_pipeline.py -m gpu -q -x -p no:cacheprovider -k "(test_small_linear_decode_products and (449 or 80)) or test_device_patch_preprocessor or (wire_batches and spectext)" > gpurun_out/r2_sanitizer3_kern.log 2>&1
echo "small_linear row blocks + derivative patches + spectext collate: memcheck rc=$?" | tee -a $S
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer3_kern.log | tail -2 | tee -a $S
