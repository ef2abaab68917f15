"""COO -> CSC of matrices whose entries come row by row (convert.cu, transpose_entries) against the radix sort.
  python scripts/transpose_bench.py [n_stencil]      (default 256: configs[1], 449 M entries)
THSP_NO_TRANSPOSE=1 in the environment sends the same calls through the radix sort (the before number)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H
from arm_spmv_b200.lib import load

lib = load()
torch.cuda.set_device(0)


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def row(name, A):
    nnz = A.nnz
    ms = timeit(lambda: H.CSCMatrix(A))
    path = lib.thsp_coo_last_path()
    minimum = nnz * 16 + nnz * 12 + (A.ncol + 1) * 4
    print(f"{name}: COO->CSC {ms:8.3f} ms  path {path}  {nnz / ms / 1e6:6.1f} G entries/s  "
          f"{minimum / ms / 1e6:7.1f} GB/s of the minimum traffic ({nnz} entries)", flush=True)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
row("lap5 1024^2 (configs[0])", H.lap5_coo(1024))
row(f"stencil27 {n}^3", H.stencil27_coo(n))
if n >= 256 and len(sys.argv) <= 2:
    # configs[3]'s matrix ordered by row (what a CSR-written file of it looks like): random columns, no locality
    A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43)
    o = torch.sort(A.row_ind.long() * (1 << 23) + A.col_ind.long()).indices
    B = H.COOMatrix(A.nrow, A.ncol, A.row_ind[o].contiguous(), A.col_ind[o].contiguous(), A.values[o].contiguous())
    del A, o
    row("uniform 8M x 8M ordered by row", B)
