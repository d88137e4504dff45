"""GEMM diagnostics: the C2 shapes under MMA_GEMM_DBG (bit0: epilogue without global traffic, bit1: no loads / MMAs,
bit2: loads but no MMAs).  Run once per MMA_GEMM_DBG value; results are only meaningful as time, not numerics."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.gemm_bench import run  # noqa: E402
from multimodalanalytical_b200._lib import EPI_DGELU, EPI_GELU, EPI_RESID, EPI_STORE  # noqa: E402

print("MMA_GEMM_DBG =", os.environ.get("MMA_GEMM_DBG"), flush=True)
run("tall K=4096 (M16384,N2048)", 16384, 2048, 4096, EPI_STORE, bias=False)
run("qkv store", 16384, 1536, 512, EPI_STORE)
run("qkv store max_ctas=37", 16384, 1536, 512, EPI_STORE, max_ctas=37)
run("ffn1 gelu", 16384, 2048, 512, EPI_GELU)
run("out-proj resid f32", 16384, 512, 512, EPI_RESID, out_dt=torch.float32)
run("ffn2 resid f32", 16384, 512, 2048, EPI_RESID, out_dt=torch.float32)
run("dgrad ffn2 dgelu (B mn)", 16384, 2048, 512, EPI_DGELU, b_mn=True)
