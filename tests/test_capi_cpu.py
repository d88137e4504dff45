"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports exactly the
symbols include/mma_b200.h declares; the ctypes mirror of `struct Epi` matches the C layout; product code
refuses CPU tensors loudly (no fallback)."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest
import torch

from multimodalanalytical_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mma_b200.h")


def header_functions():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^int\s+(mma_\w+)\s*\(", src, flags=re.M)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _lib.build()
    lib = ctypes.CDLL(path)
    decl = header_functions()
    assert decl, "no declarations parsed"
    assert sorted(_lib.EXPORTS) == decl
    for name in decl:
        assert hasattr(lib, name), name


def test_sass_contains_tcgen05_and_tma():
    out = subprocess.run(["cuobjdump", "-sass", _lib.build()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in out, "TMA loads missing from SASS"
    assert "LDTM" in out, "tcgen05.ld missing from SASS"


def test_epi_struct_layout_matches_c():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "mma_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(Epi), offsetof(Epi, out), offsetof(Epi, ldo), offsetof(Epi, p_drop),
         offsetof(Epi, seed), offsetof(Epi, drop_ld));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", c, "-o", exe],
                       check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    E = _lib.Epi
    want = [ctypes.sizeof(E), E.out.offset, E.ldo.offset, E.p_drop.offset, E.seed.offset, E.drop_ld.offset]
    assert got == want


def test_decode_step_struct_layout_matches_c():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "mma_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(MmaDecodeLayer), sizeof(MmaDecodeStep), offsetof(MmaDecodeStep, tok),
         offsetof(MmaDecodeStep, logits), offsetof(MmaDecodeStep, ldv), offsetof(MmaDecodeStep, gated),
         offsetof(MmaDecodeStep, scale));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", c, "-o", exe],
                       check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    L, S = _lib.DecodeLayer, _lib.DecodeStep
    want = [ctypes.sizeof(L), ctypes.sizeof(S), S.tok.offset, S.logits.offset, S.ldv.offset, S.gated.offset, S.scale.offset]
    assert got == want
    assert ctypes.sizeof(S) < 4096  # passed by value as a kernel parameter


def test_epilogue_enum_matches_header():
    src = open(HEADER).read()
    for name in ("EPI_STORE", "EPI_GELU", "EPI_RESID", "EPI_DGELU", "EPI_GLU_MUL", "EPI_DGLU", "EPI_ACCUM", "EPI_RELU",
                 "EPI_DRELU"):
        m = re.search(rf"{name}\s*=\s*(\d+)", src)
        assert m and int(m.group(1)) == getattr(_lib, name), name


def test_ops_refuse_cpu_tensors():
    x = torch.zeros(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ln_fwd(x, None, None, x.clone())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(x, x, 4, 4, 8, _lib.Epi())
