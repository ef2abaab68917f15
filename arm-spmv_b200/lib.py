"""ctypes binding of libthsparse_cuda.so (the C ABI declared in include/thsp.h).

There is no fallback: if the shared library is missing ``load()`` raises, and every compute
entry point returns an error when no CUDA device is usable."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libthsparse_cuda.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "thsp.h")

CSR_AUTO, CSR_SCALAR, CSR_VECTOR, CSR_STREAM, CSR_MERGE = 0, 1, 2, 3, 4
KERNEL_NAMES = {0: "auto", 1: "scalar", 2: "vector", 3: "stream", 4: "merge"}


class ThspError(RuntimeError):
    pass


_lib = None


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into csrc/libthsparse_cuda.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise ThspError("building libthsparse_cuda.so failed")
    return SO_PATH


def declared_symbols() -> list[str]:
    """Every function include/thsp.h declares with THSP_API."""
    text = open(HEADER).read()
    return sorted(set(re.findall(r"THSP_API[^;(]*?\b(thsp_\w+)\s*\(", text)))


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ThspError(f"{SO_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        lib = C.CDLL(SO_PATH)
        lib.thsp_last_error.restype = C.c_char_p
        lib.thsp_version.restype = C.c_char_p
        lib.thsp_launch_count.restype = C.c_uint64
        lib.thsp_stencil27_nnz.restype = C.c_int64
        lib.thsp_stencil27_nnz.argtypes = [C.c_int, C.c_int64, C.c_int64]
        lib.thsp_lap5_nnz.restype = C.c_int64
        lib.thsp_cg_work_doubles.restype = C.c_int64
        lib.thsp_cg_work_doubles.argtypes = [C.c_int64]
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise ThspError(load().thsp_last_error().decode(errors="replace") or f"thsp call failed ({rc})")


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(int(t))


def current_stream() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().thsp_launch_count())
