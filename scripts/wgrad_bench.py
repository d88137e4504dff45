"""Grouped weight-gradient kernel on one decoder layer's / one encoder layer's products (C2 shapes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalanalytical_b200 import ops  # noqa: E402
from scripts.gemm_bench import timeit  # noqa: E402

dev = "cuda"


def layer(R, Re, dec=True):
    d, f = 512, 2048
    specs = [(R, 3 * d, d), (R, d, d), (R, f, d), (R, d, f)]
    if dec:
        specs += [(R, d, d), (Re, 2 * d, d), (R, d, d)]
    items, flops = [], 0.0
    for (r, n, k) in specs:
        dy = torch.randn(r, n, device=dev).to(torch.bfloat16)
        x = torch.randn(r, k, device=dev).to(torch.bfloat16)
        items.append((dy, x, torch.zeros(n, k, device=dev), torch.zeros(n, device=dev), n, k, r))
        flops += 2.0 * r * n * k
    return items, flops


print("MMA_WGRAD2 =", os.environ.get("MMA_WGRAD2"))
for name, (items, fl) in (("decoder layer", layer(16384, 9216, True)), ("encoder layer", layer(9216, 9216, False))):
    t = timeit(lambda: ops.wgrad_group(items), n=10, inner=5)
    print(f"{name:16s} {t:8.1f} us  {fl / t / 1e6:8.1f} TFLOP/s", flush=True)
