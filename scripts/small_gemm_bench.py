"""us per launch of the decode-shaped products (M = rows, N = 512) for the kernels that can take them.
   MMA_GEMM2_MIN_TILES=1 python scripts/small_gemm_bench.py   -> forces the CTA-pair kernel"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalanalytical_b200 import ops
from multimodalanalytical_b200._lib import EPI_RESID, EPI_STORE, EPI_GELU

def bench(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

out = {}
for M in (2560, 10240):
    for (N, K, kind) in ((512, 512, "resid"), (512, 2048, "resid"), (1536, 512, "store"), (2048, 512, "gelu"), (512, 512, "store")):
        A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
        bias = torch.randn(N, device="cuda")
        if kind == "resid":
            x = torch.randn(M, N, device="cuda"); o = torch.empty_like(x)
            epi = ops.make_epi(EPI_RESID, o, bias=bias, resid=x)
        elif kind == "gelu":
            o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); epi = ops.make_epi(EPI_GELU, o, bias=bias)
        else:
            o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); epi = ops.make_epi(EPI_STORE, o, bias=bias)
        out[f"{M}x{N}x{K}:{kind}"] = round(bench(lambda: ops.gemm(A, W, M, N, K, epi)), 2)
print(json.dumps({"min_tiles": os.environ.get("MMA_GEMM2_MIN_TILES"), "us": out}))
