import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalanalytical_b200 import ops
from multimodalanalytical_b200._lib import EPI_STORE
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
def run_dbg(tag, M, N, K, acc, reps=1):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    epi = ops.make_epi(EPI_STORE, out, accumulate=acc)
    def f():
        for _ in range(reps): ops.gemm(A, B, M, N, K, epi)
    t = timeit(f) / reps
    print(f"{tag:46s} M={M:6d} N={N:5d} K={K:5d}  {t:8.1f} us  {2.0*M*N*K/t/1e6:8.1f} TFLOP/s", flush=True)
for acc, name in ((0, "normal"), (100, "no stores"), (101, "no tmem-ld, no stores")):
    run_dbg(f"qkv {name}", 16384, 1536, 512, acc)
    run_dbg(f"qkv {name} x8 back-to-back", 16384, 1536, 512, acc, reps=8)
for acc, name in ((0, "normal"), (101, "handshake only")):
    run_dbg(f"1 tile {name}", 128, 256, 512, acc)
    run_dbg(f"1 tile x16 back-to-back {name}", 128, 256, 512, acc, reps=16)
    run_dbg(f"148 tiles (1/SM) {name}", 128 * 148, 256, 512, acc)
    run_dbg(f"296 tiles (2/SM) {name}", 128 * 296, 256, 512, acc)
    run_dbg(f"592 tiles (4/SM) {name}", 128 * 592, 256, 512, acc)
    run_dbg(f"1184 tiles (8/SM) {name}", 128 * 1184, 256, 512, acc)
    run_dbg(f"1184 tiles (8/SM) K=2048 {name}", 128 * 1184, 256, 2048, acc)
