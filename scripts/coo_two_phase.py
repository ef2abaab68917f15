"""COO SpMV on configs[3] (uniform 8M x 8M, 128 M unsorted entries) and on the R-MAT matrix: the fused kernel against
the two-phase path (products, then scatter).  THSP_COO_TWO_PHASE=0|1 forces a path, unset = the library decides.
Checks y against the CSR result row by row (1e-12 of sum |a x|)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H

torch.cuda.set_device(0)


def timeit(fn, reps=10, warm=3):
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


mode = os.environ.get("THSP_COO_TWO_PHASE", "auto")
for w in sys.argv[1:] or ["uniform", "rmat"]:
    A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43) if w == "uniform" else H.rmat_coo(24, 16 << 24, 42)
    x = H.gen_vector(A.ncol, 3)
    y = H.Vector(A.nrow); y.Fill(0.0)
    H.COOMatirxMatVector(A, x, y)
    B = H.CSRMatrix(A)
    yr = H.Vector(A.nrow); yr.Fill(0.0)
    H.CSRMatrixMatVector(B, x, yr)
    scale = torch.zeros(A.nrow, dtype=torch.float64, device="cuda")
    scale.index_add_(0, A.row_ind.to(torch.int64), (A.values * x.values[A.col_ind.to(torch.int64)]).abs())
    err = float(((y.values - yr.values).abs() / scale.clamp_min(1e-300)).max())
    del B, yr, scale
    ms = timeit(lambda: H.COOMatirxMatVector(A, x, y))
    nb = A.nnz * 16 + A.ncol * 8 + 2 * A.nrow * 8
    print(f"two_phase={mode:4s} {w:8s} COO SpMV {ms:7.3f} ms  {2.0 * A.nnz / ms / 1e6:7.1f} GFLOP/s  {nb / ms / 1e6:7.1f} GB/s  max row error {err:.2e}", flush=True)
    del A, x, y
    torch.cuda.empty_cache()
