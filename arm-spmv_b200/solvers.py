"""Callers above the SpMV path: conjugate gradients and damped Jacobi.

The reference stops at the building blocks - `CSRMatrixMatVector`, `vec_dot`, `vec_axpby` are there,
`vec_dot` / `vec_axpby` are never called (src/vec_vec.cpp:15,31) and the `diagonal` arrays are kept
"for SymGS" (include/matrix.h:36,81) without a smoother.  Both solvers below are plain compositions
of those calls on the GPU library: every vector stays in HBM, the only values that cross to the
host are the scalars of the recurrences."""
from __future__ import annotations

import ctypes as C
import math

from . import host as H
from .lib import check, current_stream, load, ptr


def csr_diagonal(A: "H.CSRMatrix") -> "H.Vector":
    d = H.Vector(A.nrow, device=A.values.device)
    check(load().thsp_csr_diagonal_f64(A.nrow, ptr(A.row_ptr), ptr(A.col_ind), ptr(A.values), ptr(d.values), current_stream()))
    return d


def residual(A, b, x, r):
    """r = b - A x   (Copy, Fill-free: y = A x then vec_axpby(1, b, -1, y, r))."""
    H.CSRMatrixMatVector(A, x, r, accumulate=False)
    H.vec_axpby(1.0, b, -1.0, r, r)


def cg(A: "H.CSRMatrix", b: "H.Vector", x: "H.Vector", tol: float = 1e-10, maxit: int = 1000):
    """Conjugate gradients for a symmetric positive definite CSR matrix.  Returns (iterations, ||r|| / ||b||, history)."""
    n = A.nrow
    dev = A.values.device
    r, p, Ap = H.Vector(n, device=dev), H.Vector(n, device=dev), H.Vector(n, device=dev)
    residual(A, b, x, r)
    p.Copy(r)
    rs = H.vec_dot(r, r)
    bnorm = math.sqrt(H.vec_dot(b, b)) or 1.0
    hist = [math.sqrt(rs) / bnorm]
    it = 0
    while it < maxit and hist[-1] > tol:
        H.CSRMatrixMatVector(A, p, Ap, accumulate=False)
        alpha = rs / H.vec_dot(p, Ap)
        x.AddScaled(alpha, p)            # x += alpha p      (Vector::AddScaled, src/vector.cpp:98-128)
        r.AddScaled(-alpha, Ap)          # r -= alpha A p
        rs_new = H.vec_dot(r, r)
        H.vec_axpby(1.0, r, rs_new / rs, p, p)   # p = r + beta p   (alpha == 1 branch, src/vec_vec.cpp:54-61)
        rs = rs_new
        it += 1
        hist.append(math.sqrt(rs) / bnorm)
    return it, hist[-1], hist


def jacobi(A: "H.CSRMatrix", b: "H.Vector", x: "H.Vector", sweeps: int, omega: float = 1.0, diag: "H.Vector | None" = None):
    """`sweeps` damped Jacobi sweeps x += omega D^-1 (b - A x).  Returns ||b - A x|| / ||b|| after the last one."""
    n = A.nrow
    d = diag if diag is not None else csr_diagonal(A)
    r = H.Vector(n, device=A.values.device)
    for _ in range(sweeps):
        residual(A, b, x, r)
        check(load().thsp_jacobi_update_f64(C.c_int64(n), C.c_double(omega), ptr(d.values), ptr(r.values), ptr(x.values), current_stream()))
    residual(A, b, x, r)
    return math.sqrt(H.vec_dot(r, r)) / (math.sqrt(H.vec_dot(b, b)) or 1.0)
