"""One launch of every shipped SpMV kernel on the matrix its BASELINE config names - the target of the round's ncu
captures (profiles/README.md).  Usage: python scripts/prof_all.py stencil|rmat|uniform"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H

torch.cuda.set_device(0)
which = sys.argv[1]
cudart = torch.cuda.cudart()


def prof(fn):
    """Only what runs inside is captured (ncu --profile-from-start off)."""
    torch.cuda.synchronize()
    cudart.cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()


if which == "stencil":      # configs[1]: 27-point stencil 256^3, every format
    n = 256
    N = n ** 3
    B = H.stencil27_csr(n)
    x = H.gen_vector(N, 3)
    y = H.Vector(N); y.Fill(0.0)
    B.plan()
    prof(lambda: H.CSRMatrixMatVector(B, x, y))                     # csr_stream_kernel
    prof(lambda: H.csr_spmv_kernel(2, 8, B, x.values, y.values))    # csr_vector_kernel<8>
    E = H.stencil27_ell(n); prof(lambda: H.ELLMatrixMatVector(E, x, y)); del E
    D = H.DIAMatrix(B); prof(lambda: H.DIAMatrixMatVector(D, x, y)); del D
    A = H.stencil27_coo(n); prof(lambda: H.COOMatirxMatVector(A, x, y))
    Cc = H.CSCMatrix(A); prof(lambda: H.CSCMatrixMatVector(Cc, x, y))
    del A, Cc
    from arm_spmv_b200 import solvers                                # SymGS: every colour of one forward + backward sweep
    S = solvers.SymGS(B)
    r = H.gen_vector(N, 5)
    S.sweep(r, y)
    prof(lambda: S.sweep(r, y))
elif which == "rmat":       # configs[2]
    A = H.rmat_coo(24, 16 << 24, 42)
    B = H.CSRMatrix(A)
    x = H.gen_vector(B.ncol, 3)
    y = H.Vector(B.nrow); y.Fill(0.0)
    B.plan()
    H.CSRMatrixMatVector(B, x, y)                                   # builds the run table
    prof(lambda: H.CSRMatrixMatVector(B, x, y))                     # plan -> merge-path
    B32 = H.CSRMatrix(nrow=B.nrow, ncol=B.ncol, row_ptr=B.row_ptr, col_ind=B.col_ind, values=B.values.to(torch.float32))
    y32 = torch.zeros(B.nrow, dtype=torch.float32, device="cuda")
    x32 = x.values.to(torch.float32)
    prof(lambda: H.csr_spmv_kernel(4, 1, B32, x32, y32))
    del B32, y32, x32
    prof(lambda: H.COOMatirxMatVector(A, x, y))                    # coo_kernel with the head-row window
    Cc = H.CSCMatrix(A); prof(lambda: H.CSCMatrixMatVector(Cc, x, y))
else:                       # configs[3]
    A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43)
    x = H.gen_vector(A.ncol, 3)
    y = H.Vector(A.nrow); y.Fill(0.0)
    B = H.CSRMatrix(A)                                # warm-up: scratch buffers allocated
    holder = {}
    prof(lambda: holder.update(B=H.CSRMatrix(A)))     # the conversion kernels
    B = holder["B"]
    B.plan()
    prof(lambda: H.CSRMatrixMatVector(B, x, y))
    prof(lambda: H.COOMatirxMatVector(A, x, y))
    Cc = H.CSCMatrix(A); prof(lambda: H.CSCMatrixMatVector(Cc, x, y)); del Cc
    D = H.ELLMatrix(A); prof(lambda: H.ELLMatrixMatVector(D, x, y))
torch.cuda.synchronize()
print("ok")
