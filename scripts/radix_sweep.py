"""COO->CSR/CSC/ELL on configs[3] (uniform 8M x 8M, 128 M entries) and COO->CSR on configs[2] (R-MAT s24) for one
variant of the radix-sort kernels (THSP_RADIX_VARIANT, read once per process), checked against torch's stable sort.
Usage: THSP_RADIX_VARIANT=k python scripts/radix_sweep.py [uniform] [rmat]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H

torch.cuda.set_device(0)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def check(A, B):
    order = torch.sort(A.row_ind.to(torch.int64), stable=True)[1]
    ok = bool((B.col_ind == A.col_ind[order]).all()) and bool((B.values == A.values[order]).all())
    cnt = torch.bincount(A.row_ind.to(torch.int64), minlength=A.nrow)
    rp = torch.zeros(A.nrow + 1, dtype=torch.int64, device="cuda"); rp[1:] = torch.cumsum(cnt, 0)
    return ok and bool((B.row_ptr.to(torch.int64) == rp).all())


v = os.environ.get("THSP_RADIX_VARIANT", "0")
for w in sys.argv[1:] or ["uniform"]:
    A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43) if w == "uniform" else H.rmat_coo(24, 16 << 24, 42)
    B = H.CSRMatrix(A)
    good = check(A, B)
    del B
    line = f"variant {v} {w:8s} bit-exact={good}  COO->CSR {timeit(lambda: H.CSRMatrix(A)):7.3f} ms"
    line += f"  COO->CSC {timeit(lambda: H.CSCMatrix(A)):7.3f} ms"
    if w == "uniform":
        line += f"  COO->ELL {timeit(lambda: H.ELLMatrix(A)):7.3f} ms"
    print(line, flush=True)
    del A
    torch.cuda.empty_cache()
