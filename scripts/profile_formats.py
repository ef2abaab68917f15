"""Run one launch of each secondary SpMV kernel on the 27-point stencil (for ncu captures)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arm_spmv_b200 import host as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
torch.cuda.set_device(0)
A = H.stencil27_coo(n)
B = H.CSRMatrix(A)
C = H.CSCMatrix(A)
x = H.gen_vector(A.ncol, 3)
y = H.Vector(A.nrow); y.Fill(0.0)
for _ in range(2):
    H.COOMatirxMatVector(A, x, y)
    H.CSCMatrixMatVector(C, x, y)
    H.csr_spmv_kernel(4, 1, B, x.values, y.values, True)
    H.csr_spmv_kernel(2, 8, B, x.values, y.values, True)
torch.cuda.synchronize()
print("ok")
