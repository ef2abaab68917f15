"""SymGS and CG behind the C ABI at full size (27-point stencil n^3, default 256): colouring time, one symmetric sweep
against its roofline (the matrix is read twice, x gathered, r read, x written twice), CG / PCG time per iteration.
  python scripts/solver_bench.py [n]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H, solvers

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = n ** 3
torch.cuda.set_device(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
A = H.stencil27_csr(n)
t0 = time.perf_counter()
S = solvers.SymGS(A)
torch.cuda.synchronize()
print(f"stencil {n}^3: {N} rows, {A.nnz} entries; colouring: {S.ncolors} colours in {S.rounds} rounds, {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
r = H.gen_vector(N, 5)
x = H.Vector(N); x.Fill(0.0)
for _ in range(2):
    S.sweep(r, x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
a.record()
for _ in range(reps):
    S.sweep(r, x)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
byt = 2 * (A.nnz * 12 + (N + 1) * 4 + N * 4 + N * 8 * 4)   # per sweep direction: matrix, row list, r, diag, x read + written
print(f"SymGS sweep (forward + backward): {ms:.3f} ms  {byt / ms / 1e6:.0f} GB/s  {byt / ms / 1e6 / PEAK * 100:.1f}% of measured HBM  "
      f"({2 * 2 * A.nnz / ms / 1e6:.0f} GFLOP/s)", flush=True)
bvec = H.gen_vector(N, 3)
for pre in ("none", "jacobi", "symgs"):
    xx = H.Vector(N); xx.Fill(0.0)
    solvers.pcg(A, bvec, xx, tol=0.0, maxit=2, precond=pre, M=S if pre == "symgs" else None, diagonal=S.diagonal)
    xx.Fill(0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it, rel = solvers.pcg(A, bvec, xx, tol=1e-9, maxit=200, precond=pre, M=S if pre == "symgs" else None, diagonal=S.diagonal)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    print(f"CG precond={pre:7s}: {it} iterations to {rel:.2e}, {dt:.1f} ms total, {dt / max(it, 1):.3f} ms / iteration", flush=True)
