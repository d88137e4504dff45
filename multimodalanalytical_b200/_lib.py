"""Build + load the C-ABI CUDA library (`lib/libmma_b200.so`, declared in include/mma_b200.h).

The library is compiled in-tree with nvcc for sm_100a only and loaded with ctypes; every entry point takes
raw device pointers, sizes and a cudaStream_t and returns an int status (0 = ok).  There is no CPU or
PyTorch fallback: if the library is missing or a launch fails, the caller gets an exception.
"""
import ctypes as C
import os
import shutil
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libmma_b200.so")
SOURCES = ["gemm_tc.cu", "gemm_tc2.cu", "gemm_glu2.cu", "gemm_simt.cu", "rowops.cu", "attention.cu", "attention_mma.cu", "attention_tc5.cu", "trainops.cu", "ddp_p2p.cu", "decode.cu", "decode_small.cu", "decode_step.cu", "align.cu", "collate.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "550"]

MMA_BF16, MMA_F32 = 0, 1
EPI_STORE, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_GLU_MUL, EPI_DGLU, EPI_ACCUM, EPI_RELU, EPI_DRELU = range(9)


class Epi(C.Structure):
    """Mirror of `struct Epi` (csrc/common.cuh, include/mma_b200.h)."""

    _fields_ = [
        ("kind", C.c_int), ("out_f32", C.c_int), ("aux_f32", C.c_int), ("resid_f32", C.c_int),
        ("out", C.c_void_p), ("out2", C.c_void_p), ("bias", C.c_void_p), ("resid", C.c_void_p),
        ("aux", C.c_void_p), ("aux2", C.c_void_p),
        ("ldo", C.c_longlong), ("ldo2", C.c_longlong), ("ldr", C.c_longlong), ("lda", C.c_longlong),
        ("lda2", C.c_longlong),
        ("p_drop", C.c_float), ("alpha", C.c_float), ("seed", C.c_ulonglong), ("site", C.c_uint),
        ("accumulate", C.c_int), ("drop_ld", C.c_longlong),
    ]


DECODE_MAX_LAYERS = 12


class DecodeLayer(C.Structure):
    """Mirror of `MmaDecodeLayer` (include/mma_b200.h, csrc/decode_step.cu)."""

    _fields_ = [(n, C.c_void_p) for n in (
        "w_qkv", "w_so", "w_cq", "w_co", "w_f1", "w_fg", "w_f2", "b_qkv", "b_so", "b_cq", "b_co", "b_f1", "b_fg", "b_f2",
        "n1g", "n1b", "n2g", "n2b", "n3g", "n3b", "kc", "vc", "kvmem")]


class DecodeStep(C.Structure):
    """Mirror of `MmaDecodeStep`."""

    _fields_ = [("layer", DecodeLayer * DECODE_MAX_LAYERS)] + [(n, C.c_void_p) for n in (
        "tok", "emb", "emb_g", "emb_b", "pos", "cur_len", "fin_g", "fin_b", "w_lm", "b_lm", "x", "xa", "xb", "qkv", "att",
        "q", "a", "logits", "anc", "enc_mask", "dbg_times")] + [("ldv", C.c_longlong)] + [(n, C.c_int) for n in (
            "layers", "R", "rows_per_cluster", "beams", "d", "f", "H", "Lmax", "S", "V", "gated")] + [
                ("eps", C.c_float), ("scale", C.c_float)]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> lib/libmma_b200.so (sm_100a, -lineinfo).  Cross-compiles without a GPU.
    Safe under concurrent callers (one process per GPU under torchrun): the staleness check and the build run under an
    exclusive file lock, so one rank compiles and the others wait and then find the library fresh."""
    if not force and not _stale():
        return LIBPATH
    import fcntl

    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIBPATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose=False):
    objdir = os.path.join(LIBDIR, f"obj.{os.getpid()}")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = f"{LIBPATH}.{os.getpid()}.tmp"
    r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIBPATH)
    shutil.rmtree(objdir, ignore_errors=True)
    return LIBPATH


_lock = threading.Lock()
_lib = None

_vp, _i, _ll, _f, _ull, _u = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong, C.c_uint
_SIGS = {
    "mma_gemm_bf16": [_vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, C.POINTER(Epi), _i, _i, _vp],
    "mma_gemm2_resid_ln": [_vp, _ll, _vp, _ll, _i, _i, _i, C.POINTER(Epi), _vp, _vp, _f, _vp, _ll, _vp],
    "mma_gemm2_dual": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, C.POINTER(Epi), _vp],
    "mma_ffn_glu_fwd": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _f, _ull, _u, _vp],
    "mma_ffn_dglu": [_vp, _ll, _vp, _ll, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _f, _ull, _u, _ll, _vp],
    "mma_wgrad_group": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mma_gemm_simt": [_vp, _i, _ll, _ll, _vp, _i, _ll, _ll, _i, _i, _i, C.POINTER(Epi), _i, _vp],
    "mma_gather_rows": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "mma_scatter_add_rows": [_vp, _vp, _vp, _vp, _i, _i, _ll, _vp],
    "mma_ln_fwd": [_vp, _i, _ll, _vp, _vp, _f, _vp, _i, _ll, _vp, _i, _ll, _vp, _ll, _i, _i, _i, _i, _i, _vp],
    "mma_ln_bwd": [_vp, _i, _ll, _i, _i, _i, _vp, _i, _ll, _vp, _f, _vp, _ll, _vp, _ll, _vp, _i, _ll, _f, _ull, _u,
                   _vp, _vp, _i, _i, _vp],
    "mma_colsum": [_vp, _i, _ll, _vp, _i, _i, _vp],
    "mma_cast_f32_bf16": [_vp, _vp, _ll, _vp],
    "mma_cast_bf16_f32": [_vp, _vp, _ll, _vp],
    "mma_patchify": [_vp, _ll, _i, _f, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "mma_patchify_rows": [_vp, _ll, _vp, _i, _f, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "mma_patchify_deriv": [_vp, _ll, _vp, _i, _i, _f, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "mma_collate_tokens": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "mma_collate_target": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "mma_collate_values": [_vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp],
    "mma_masked_mean_fwd": [_vp, _i, _ll, _vp, _vp, _i, _i, _i, _vp],
    "mma_masked_mean_bwd": [_vp, _vp, _vp, _ll, _i, _i, _i, _vp],
    "mma_align_loss": [_vp, _ll, _vp, _ll, _i, _i, _i, _f, _vp, _vp, _vp, _ll, _f, _vp],
    "mma_add_strided": [_vp, _ll, _vp, _ll, _vp],
    "mma_attn_fwd": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _ull, _u, _i,
                     _vp],
    "mma_attn_fwd_tc": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _i, _i, _i, _i, _i, _f, _f, _ull, _u, _vp],
    "mma_attn_bwd_tc": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll,
                        _i, _i, _i, _i, _i, _f, _f, _ull, _u, _vp],
    "mma_attn_fwd_t5": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _i, _i, _i, _i, _i, _f, _f, _ull, _u, _vp],
    "mma_attn_bwd_t5": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll,
                        _i, _i, _i, _i, _i, _f, _f, _ull, _u, _vp],
    "mma_attn_fwd_t5b": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _ull, _u,
                         _vp],
    "mma_attn_bwd_t5b": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp,
                         _vp, _i, _i, _i, _i, _i, _f, _f, _ull, _u, _vp],
    "mma_attn_bwd": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i,
                     _i, _i, _i, _i, _f, _f, _ull, _u, _i, _vp],
    "mma_ce_fwd": [_vp, _ll, _vp, _i, _i, _f, _ll, _vp, _vp, _vp, _vp],
    "mma_ce_bwd": [_vp, _ll, _vp, _vp, _vp, _f, _i, _i, _f, _ll, _vp, _i, _ll, _vp],
    "mma_grad_norm": [_vp, _ll, _vp, _vp, _vp],
    "mma_adam_step": [_vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _i, _i, _vp],
    "mma_add_u64": [_vp, _ull, _vp],
    "mma_p2p_barrier": [_vp, _vp, _i, _i, _vp],
    "mma_p2p_reduce_shard": [_vp, _vp, _i, _i, _ll, _ll, _vp, _vp, _vp],
    "mma_p2p_adam_shard": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _ll, _ll, _vp, _i, _vp],
    "mma_decode_embed": [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp],
    "mma_small_linear": [_vp, _i, _ll, _vp, _vp, _f, _vp, _vp, _ll, _vp, _vp, _vp, _ll, _vp, _i, _ll, _i, _i, _i, _i, _vp],
    "mma_decode_step": [_vp, _i, _vp],
    "mma_decode_step_max_clusters": [_i],
    "mma_decode_self_attn": [_vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _f, _i, _i, _vp],
    "mma_decode_cross_attn": [_vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _f, _i, _vp],
    "mma_beam_step": [_vp, _ll, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                      _vp, _vp],
    "mma_greedy_step": [_vp, _ll, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "mma_advance": [_vp, _vp],
    "mma_beam_step_ex": [_vp, _ll, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                         _vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "mma_greedy_step_ex": [_vp, _ll, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "mma_score_rows": [_vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _i, _vp],
    "mma_guided_mask": [_vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp],
}
EXPORTS = tuple(_SIGS)


def load():
    """ctypes handle to the library; builds it if the in-tree .so is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            # MMA_B200_LIB: load a prebuilt library instead (A/B measurements of two builds on one box)
            path = os.environ.get("MMA_B200_LIB") or build()
            lib = C.CDLL(path)
            for name, sig in _SIGS.items():
                fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
                fn.argtypes = sig
                fn.restype = C.c_int
            _lib = lib
    return _lib


class KernelError(RuntimeError):
    pass


_ERRS = {-1: "bad argument", -2: "kernel launch failed", -3: "unsupported shape/type", -4: "driver / tensor-map error"}


def check(rc, what):
    if rc != 0:
        raise KernelError(f"{what}: {_ERRS.get(rc, 'error')} (status {rc})")
