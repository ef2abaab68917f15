"""Time the merge-path CSR kernel on the R-MAT matrix (BASELINE configs[2]); THSP_MERGE_VARIANT picks the build variant."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H

torch.cuda.set_device(0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
A = H.rmat_coo(scale, 16 << scale, 42)
t0 = time.perf_counter(); B = H.CSRMatrix(A); torch.cuda.synchronize()
print(f"COO->CSR {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
del A
x = H.gen_vector(B.ncol, 3)
y = H.Vector(B.nrow); y.Fill(0.0)


def timed(tag, fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    nb = B.nnz * 12 + (B.nrow + 1) * 4 + B.ncol * 8 + 2 * B.nrow * 8
    print(f"{tag:28s} {ms:8.4f} ms  {2 * B.nnz / ms / 1e6:7.1f} GFLOP/s  {nb / ms / 1e6:7.1f} GB/s", flush=True)


timed(f"merge variant {os.environ.get('THSP_MERGE_VARIANT', '0')}", lambda: H.csr_spmv_kernel(4, 1, B, x.values, y.values, True))
if os.environ.get("THSP_MERGE_VARIANT", "0") == "0":
    timed("vector32", lambda: H.csr_spmv_kernel(2, 32, B, x.values, y.values, True))
    B32 = H.CSRMatrix(nrow=B.nrow, ncol=B.ncol, row_ptr=B.row_ptr, col_ind=B.col_ind, values=B.values.to(torch.float32))
    x32 = x.values.to(torch.float32); y32 = torch.zeros(B.nrow, dtype=torch.float32, device="cuda")
    timed("merge fp32", lambda: H.csr_spmv_kernel(4, 1, B32, x32, y32, True))
