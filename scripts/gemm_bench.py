"""Micro-benchmark of the tcgen05 GEMM on the shapes of the C2 training step (CUDA events, median of 20)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalanalytical_b200 import ops  # noqa: E402
from multimodalanalytical_b200._lib import EPI_ACCUM, EPI_DGELU, EPI_GELU, EPI_RESID, EPI_STORE  # noqa: E402

dev = "cuda"


def timeit_graph(fn, n=10, inner=20):
    """us per launch with `inner` launches captured in ONE CUDA graph (no host launch path at all: what the kernel
    costs inside the captured training step, PDL overlap included)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort()
    return ts[len(ts) // 2]


def timeit(fn, n=10, inner=10):
    if os.environ.get("BENCH_GRAPH", "1") != "0":
        return timeit_graph(fn, n=n)
    """us per launch: `inner` back-to-back launches per event pair (the GPU queue never drains, so the Python /
    launch path is not part of the measurement), median over `n` repeats."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort()
    return ts[len(ts) // 2]


def run(tag, M, N, K, kind, out_dt=torch.bfloat16, b_mn=False, max_ctas=0, bias=True):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=out_dt)
    kw = {}
    if bias and kind in (EPI_STORE, EPI_GELU, EPI_RESID):
        kw["bias"] = torch.randn(N, device=dev)
    if kind == EPI_GELU:
        kw["out2"] = torch.empty_like(out)
        kw.update(p_drop=0.1, seed=1, site=3)
    if kind == EPI_RESID:
        kw["resid"] = torch.randn(M, N, device=dev)
        kw.update(p_drop=0.1, seed=1, site=3)
    if kind == EPI_DGELU:
        kw["aux"] = torch.randn(M, N, device=dev).to(torch.bfloat16)
        kw.update(p_drop=0.1, seed=1, site=3, drop_ld=N)
    epi = ops.make_epi(kind, out, **kw)
    t = timeit(lambda: ops.gemm(A, B, M, N, K, epi, b_mn=b_mn, max_ctas=max_ctas))
    print(f"{tag:42s} M={M:6d} N={N:5d} K={K:5d}  {t:8.1f} us  {2.0 * M * N * K / t / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print("MMA_GEMM_BN =", os.environ.get("MMA_GEMM_BN"))
    run("square 8192^3 store bf16", 8192, 8192, 8192, EPI_STORE, bias=False)
    run("square 4096x4096x4096", 4096, 4096, 4096, EPI_STORE, bias=False)
    run("tall K=4096 (M16384,N2048)", 16384, 2048, 4096, EPI_STORE, bias=False)
    run("qkv store", 16384, 1536, 512, EPI_STORE)
    run("qkv store no-bias", 16384, 1536, 512, EPI_STORE, bias=False)
    run("q store", 16384, 512, 512, EPI_STORE)
    run("out-proj resid f32", 16384, 512, 512, EPI_RESID, out_dt=torch.float32)
    run("ffn1 gelu", 16384, 2048, 512, EPI_GELU)
    run("ffn1 store only", 16384, 2048, 512, EPI_STORE)
    run("ffn2 resid f32", 16384, 512, 2048, EPI_RESID, out_dt=torch.float32)
    run("dgrad ffn2 dgelu (B mn)", 16384, 2048, 512, EPI_DGELU, b_mn=True)
    run("dgrad ffn1 store (B mn)", 16384, 512, 2048, EPI_STORE, b_mn=True, bias=False)
    run("dgrad qkv store (B mn)", 16384, 512, 1536, EPI_STORE, b_mn=True, bias=False)
    for mc in (37, 74, 111, 148):
        run(f"qkv store max_ctas={mc}", 16384, 1536, 512, EPI_STORE, max_ctas=mc)
