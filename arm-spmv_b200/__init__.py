"""B200-native SpMV behind the arm-spmv API.

Layout
  csrc/            CUDA kernels + the C ABI (include/thsp.h)  -> csrc/libthsparse_cuda.so
  csrc/host/       the reference's C++ classes re-implemented as callers of the C ABI -> bin/TH_sparse.a
  lib.py           ctypes binding of the C ABI (fails loudly when the .so is missing)
  host.py          Python mirror of the reference's operator interface on torch CUDA tensors
  power.py         row-partitioned power iteration (one process per GPU, torch.distributed)
"""
from . import lib  # noqa: F401
from .lib import build, load, ThspError  # noqa: F401
