"""Launch each kernel of BASELINE configs[2] (R-MAT s24) and configs[3] (uniform 8M) a few times - meant to run
under ncu (kernel-name filter on the command line) and, without ncu, to print CUDA-event timings.
Usage: python scripts/prof_c3c4.py [rmat|uniform] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H
from arm_spmv_b200.lib import check, current_stream, load, ptr

lib = load()
torch.cuda.set_device(0)
which = sys.argv[1] if len(sys.argv) > 1 else "rmat"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1


def timed(tag, fn):
    fn()
    if os.environ.get("PROF"):   # under ncu: one launch per kernel is enough
        torch.cuda.synchronize()
        return
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{tag:32s} {a.elapsed_time(b) / reps:9.4f} ms", flush=True)


if which == "rmat":
    A = H.rmat_coo(24, 16 << 24, 42)
else:
    A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43)
x = H.gen_vector(A.ncol, 3)
y = H.Vector(A.nrow)
y.Fill(0.0)
timed("COO->CSR", lambda: H.CSRMatrix(A))
B = H.CSRMatrix(A)
print("plan picks", B.plan_kernel(), flush=True)
timed("CSR plan kernel", lambda: H.CSRMatrixMatVector(B, x, y))
timed("CSR merge", lambda: H.csr_spmv_kernel(4, 1, B, x.values, y.values, True))
timed("CSR vector32", lambda: H.csr_spmv_kernel(2, 32, B, x.values, y.values, True))
timed("CSR vector8", lambda: H.csr_spmv_kernel(2, 8, B, x.values, y.values, True))
if which != "rmat":
    timed("CSR stream", lambda: H.csr_spmv_kernel(3, 1, B, x.values, y.values, True))
timed("COO", lambda: H.COOMatirxMatVector(A, x, y))
timed("COO->CSC", lambda: H.CSCMatrix(A))
Cc = H.CSCMatrix(A)
timed("CSC", lambda: H.CSCMatrixMatVector(Cc, x, y))
del Cc
if which != "rmat":
    timed("COO->ELL", lambda: H.ELLMatrix(A))
print("ok")
