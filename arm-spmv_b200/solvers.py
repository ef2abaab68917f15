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


class SymGS:
    """Symmetric Gauss-Seidel on the GPU (thsp_symgs_*): the smoother the reference's `diagonal` arrays were kept for
    (include/matrix.h:36,81).  The plan colours the rows once; sweep() does forward + backward, colour by colour."""

    def __init__(self, A: "H.CSRMatrix", diagonal: "H.Vector | None" = None, snapshot: bool = True):
        """snapshot: keep a copy of the matrix permuted by colour, so that colours are streamed (thsp_symgs_plan_create)."""
        self.A = A
        self.diagonal = diagonal if diagonal is not None else csr_diagonal(A)
        h = C.c_void_p()
        check(load().thsp_symgs_plan_create(C.byref(h), A.nrow, ptr(A.row_ptr), ptr(A.col_ind), ptr(A.values) if snapshot else None,
                                            current_stream()))
        self.plan = h
        st = C.c_int(0)
        check(load().thsp_symgs_plan_streams(h, C.byref(st)))
        self.streams = bool(st.value)
        nc, rounds = C.c_int(), C.c_int()
        check(load().thsp_symgs_plan_info(h, C.byref(nc), C.byref(rounds), None, 0, None, None))
        self.ncolors, self.rounds = nc.value, rounds.value

    def coloring(self):
        """(color_ptr [ncolors + 1], perm, color) as numpy arrays (tests: validity of the colouring, the oracle's sweep)."""
        import numpy as np
        import torch
        cp = (C.c_int * (self.ncolors + 1))()
        perm, color = C.c_void_p(), C.c_void_p()
        check(load().thsp_symgs_plan_info(self.plan, None, None, cp, self.ncolors + 1, C.byref(perm), C.byref(color)))
        n = self.A.nrow
        out = torch.empty(2 * n, dtype=torch.int32, device=self.A.values.device)
        check(load().thsp_memcpy_d2d(ptr(out), perm, C.c_size_t(4 * n), current_stream()))
        check(load().thsp_memcpy_d2d(C.c_void_p(out.data_ptr() + 4 * n), color, C.c_size_t(4 * n), current_stream()))
        torch.cuda.synchronize()
        h = out.cpu().numpy()
        return np.array(list(cp), dtype=np.int32), h[:n].copy(), h[n:].copy()

    def sweep(self, r: "H.Vector", x: "H.Vector"):
        A = self.A
        check(load().thsp_symgs_f64(self.plan, A.nrow, ptr(A.row_ptr), ptr(A.col_ind), ptr(A.values), ptr(self.diagonal.values),
                                    ptr(r.values), ptr(x.values), current_stream()))

    def __del__(self):
        try:
            load().thsp_symgs_plan_destroy(self.plan)
        except Exception:
            pass


def pcg(A: "H.CSRMatrix", b: "H.Vector", x: "H.Vector", tol: float = 1e-10, maxit: int = 1000, precond: str = "none",
        M: "SymGS | None" = None, diagonal: "H.Vector | None" = None):
    """thsp_cg_f64: the whole loop inside the library (scalars on the device, canonical dots).  precond: none | jacobi |
    symgs.  Returns (iterations, ||r|| / ||b||)."""
    import torch
    kind = {"none": 0, "jacobi": 1, "symgs": 2}[precond]
    if kind == 2 and M is None:
        M = SymGS(A, diagonal)
    d = M.diagonal if M is not None else (diagonal if diagonal is not None else (csr_diagonal(A) if kind else None))
    n = A.nrow
    work = torch.empty(int(load().thsp_cg_work_doubles(n)), dtype=torch.float64, device=A.values.device)
    it, rel = C.c_int(0), C.c_double(0.0)
    check(load().thsp_cg_f64(A.plan(), n, kind, M.plan if M is not None else None, ptr(A.row_ptr), ptr(A.col_ind), ptr(A.values),
                             ptr(d.values) if d is not None else None, ptr(b.values), ptr(x.values), maxit, C.c_double(tol), ptr(work),
                             C.byref(it), C.byref(rel), current_stream()))
    return it.value, rel.value


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
