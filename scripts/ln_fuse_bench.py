"""Fused residual-GEMM + LayerNorm against the unfused pair of launches (CUDA-graph timing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalanalytical_b200 import ops  # noqa: E402
from multimodalanalytical_b200._lib import EPI_RESID  # noqa: E402
from scripts.gemm_bench import timeit  # noqa: E402

dev = "cuda"
for M, K in ((16384, 512), (16384, 2048), (9216, 512), (9216, 2048)):
    N = 512
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias, gamma, beta = torch.randn(N, device=dev), torch.ones(N, device=dev), torch.zeros(N, device=dev)
    resid = torch.randn(M, N, device=dev)
    x = torch.empty(M, N, device=dev)
    h = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    epi = ops.make_epi(EPI_RESID, x, bias=bias, resid=resid, p_drop=0.1, seed=1, site=3)
    t_f = timeit(lambda: ops.gemm_resid_ln(A, W, M, N, K, epi, gamma, beta, h))
    t_g = timeit(lambda: ops.gemm(A, W, M, N, K, epi))
    t_l = timeit(lambda: ops.ln_fwd(x, gamma, beta, h))

    def both():
        ops.gemm(A, W, M, N, K, epi)
        ops.ln_fwd(x, gamma, beta, h)
    t_b = timeit(both)
    print(f"M={M} K={K}: fused {t_f:.1f} us | gemm {t_g:.1f} + ln {t_l:.1f} = back-to-back pair {t_b:.1f} us", flush=True)
