"""Time the other BASELINE.json configurations at full size (one GPU): every format's SpMV and the
GPU conversions, with algorithmic bytes over CUDA-event time against the measured
HBM peak.  Usage: python scripts/bench_configs.py [lap5|rmat|uniform|stencil] ..."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H
from arm_spmv_b200.lib import check, current_stream, load, ptr

lib = load()
torch.cuda.set_device(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=20, warm=3):
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


def report(tag, ms, nbytes, flops=None):
    gbs = nbytes / ms / 1e6
    s = f"{tag:42s} {ms:9.4f} ms  {gbs:8.1f} GB/s  {gbs / PEAK * 100:6.1f}% of measured"
    if flops:
        s += f"  {flops / ms / 1e6:8.1f} GFLOP/s"
    print(s, flush=True)


def spmv_all(A, name, ell=True, dia=False, kernels=()):
    nrow, ncol, nnz = A.nrow, A.ncol, A.nnz
    x = H.gen_vector(ncol, 3)
    y = H.Vector(nrow); y.Fill(0.0)
    fl = 2.0 * nnz
    t0 = time.perf_counter(); B = H.CSRMatrix(A); torch.cuda.synchronize(); t_csr = (time.perf_counter() - t0) * 1e3
    report(f"{name}: COO->CSR (first call)", t_csr, nnz * 16 + nnz * 12 + (nrow + 1) * 4)
    ms = timeit(lambda: H.CSRMatrix(A), reps=3, warm=1)
    report(f"{name}: COO->CSR", ms, nnz * 16 + nnz * 12 + (nrow + 1) * 4)
    ms = timeit(lambda: H.CSCMatrix(A), reps=3, warm=1)
    report(f"{name}: COO->CSC", ms, nnz * 16 + nnz * 12 + (ncol + 1) * 4)
    csr_b = nnz * 12 + (nrow + 1) * 4 + ncol * 8 + 2 * nrow * 8
    print(f"   plan picks: {B.plan_kernel()}")
    ms = timeit(lambda: H.CSRMatrixMatVector(B, x, y)); report(f"{name}: CSR auto {B.plan_kernel()}", ms, csr_b, fl)
    check(lib.thsp_csr_plan_autotune(B.plan(), current_stream()))
    ms = timeit(lambda: H.CSRMatrixMatVector(B, x, y)); report(f"{name}: CSR autotuned {B.plan_kernel()}", ms, csr_b, fl)
    for kid, lanes, kn in kernels:
        ms = timeit(lambda: H.csr_spmv_kernel(kid, lanes, B, x.values, y.values, True)); report(f"{name}: CSR {kn}", ms, csr_b, fl)
    ms = timeit(lambda: H.COOMatirxMatVector(A, x, y)); report(f"{name}: COO", ms, nnz * 16 + ncol * 8 + 2 * nrow * 8, fl)
    Cc = H.CSCMatrix(A)
    ms = timeit(lambda: H.CSCMatrixMatVector(Cc, x, y)); report(f"{name}: CSC", ms, nnz * 12 + (ncol + 1) * 4 + ncol * 8 + 2 * nrow * 8, fl)
    del Cc
    if ell:
        ms = timeit(lambda: H.ELLMatrix(A), reps=3, warm=1)
        D = H.ELLMatrix(A)
        K = D.nonzeros_in_row
        report(f"{name}: COO->ELL (K={K})", ms, nnz * 16 + nrow * K * 12)
        ms = timeit(lambda: H.ELLMatrixMatVector(D, x, y)); report(f"{name}: ELL", ms, nrow * K * 12 + ncol * 8 + 2 * nrow * 8, fl)
        del D
    if dia:
        ms = timeit(lambda: H.DIAMatrix(B), reps=3, warm=1)
        E = H.DIAMatrix(B)
        report(f"{name}: CSR->DIA (ndiags={E.ndiags})", ms, nnz * 12 + (nrow + 1) * 4 + nrow * E.ndiags * 8)
        ms = timeit(lambda: H.DIAMatrixMatVector(E, x, y)); report(f"{name}: DIA (ndiags={E.ndiags})", ms, nrow * E.ndiags * 8 + ncol * 8 + 2 * nrow * 8, fl)
    if B.values.dtype == torch.float64:
        B32 = H.CSRMatrix(nrow=nrow, ncol=ncol, row_ptr=B.row_ptr, col_ind=B.col_ind, values=B.values.to(torch.float32))
        x32 = x.values.to(torch.float32); y32 = torch.zeros(nrow, dtype=torch.float32, device="cuda")
        fn = lambda: check(lib.thsp_csr_plan_spmv_f32(B32.plan(), ptr(x32), ptr(y32), 1, current_stream()))
        ms = timeit(fn); report(f"{name}: CSR fp32 {B32.plan_kernel()}", ms, nnz * 8 + (nrow + 1) * 4 + ncol * 4 + 2 * nrow * 4, fl)


which = sys.argv[1:] or ["lap5", "rmat", "uniform"]
V, ST, MG = 2, 3, 4
for w in which:
    torch.cuda.empty_cache()
    if w == "lap5":      # configs[0]: 5-pt Laplacian 1024^2
        spmv_all(H.lap5_coo(1024), "lap5 1024^2", dia=True, kernels=[(1, 1, "scalar"), (V, 2, "vector2"), (V, 4, "vector4"), (ST, 1, "stream"), (MG, 1, "merge")])
    elif w == "rmat":    # configs[2]: R-MAT scale 24, avg degree 16
        spmv_all(H.rmat_coo(24, 16 << 24, 42), "rmat s24", ell=False, kernels=[(V, 8, "vector8"), (V, 16, "vector16"), (V, 32, "vector32"), (ST, 1, "stream"), (MG, 1, "merge")])
    elif w == "uniform":  # configs[3]: uniform 8M x 8M, 128M entries
        spmv_all(H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43), "uniform 8M", ell=True, kernels=[(V, 8, "vector8"), (V, 16, "vector16"), (ST, 1, "stream"), (MG, 1, "merge")])
    elif w == "lap5big":
        spmv_all(H.lap5_coo(4096), "lap5 4096^2", dia=True, kernels=[(1, 1, "scalar"), (V, 2, "vector2"), (V, 4, "vector4"), (ST, 1, "stream")])
    elif w == "stencil":
        spmv_all(H.stencil27_coo(256), "stencil 256^3", dia=True, kernels=[(V, 8, "vector8"), (V, 16, "vector16"), (V, 32, "vector32"), (MG, 1, "merge")])
