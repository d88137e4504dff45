#!/bin/bash
# which part of a pair-GEMM launch bounds it: MMA_GEMM_DBG bit0 = epilogue without global traffic, bit1 = no operand
# loads / MMAs, bit2 = operand loads but no MMAs
for d in 0 1 2 3 4; do MMA_GEMM_DBG=$d python scripts/gemm_diag.py 2>&1 | grep -v "max_ctas\|tall"; done
