"""-m gpu: the callers above the path (arm-spmv_b200/solvers.py): CG and Jacobi composed from the library's SpMV, dot and
axpby, against a dense solve of the same system on the CPU (numpy, checker only)."""
import numpy as np
import pytest

from gpu_util import dev, host

pytestmark = pytest.mark.gpu


def _dense(nrow, rp, ci, va):
    M = np.zeros((nrow, nrow))
    for r in range(nrow):
        for p in range(rp[r], rp[r + 1]):
            M[r, ci[p]] += va[p]
    return M


def test_cg_solves_the_stencil_system(thsp, cuda, oracle):
    from arm_spmv_b200 import host as H, solvers
    n = 9
    N = n ** 3
    rp, ci, va = oracle.gen_stencil27_csr(n)          # 26 on the diagonal, -1 elsewhere: symmetric, diagonally dominant at the boundary
    A = H.CSRMatrix(nrow=N, ncol=N, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    b = H.Vector(oracle.gen_vector(N, 3))
    x = H.Vector(np.zeros(N))
    it, rel, hist = solvers.cg(A, b, x, tol=1e-12, maxit=500)
    assert rel <= 1e-12 and it < 200
    want = np.linalg.solve(_dense(N, rp, ci, va), host(b.values))
    assert np.max(np.abs(host(x.values) - want)) <= 1e-9 * np.max(np.abs(want))
    assert all(h2 <= h1 * 10 for h1, h2 in zip(hist, hist[1:]))   # no blow-up on the way


def test_jacobi_and_diagonal(thsp, cuda, oracle):
    from arm_spmv_b200 import host as H, solvers
    ri, cj, v = oracle.gen_lap5_coo(20)
    N = 400
    rp, ci, va, _ = oracle.coo2csr(N, N, ri, cj, v)
    A = H.CSRMatrix(nrow=N, ncol=N, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    d = solvers.csr_diagonal(A)
    assert np.array_equal(host(d.values), np.diag(_dense(N, rp, ci, va)))
    b = H.Vector(oracle.gen_vector(N, 5))
    x = H.Vector(np.zeros(N))
    r0 = solvers.jacobi(A, b, x, 0)
    r1 = solvers.jacobi(A, b, x, 50, omega=0.8)
    r2 = solvers.jacobi(A, b, x, 200, omega=0.8)
    assert r0 == pytest.approx(1.0) and r1 < 0.9 and r2 < r1
    # one sweep against numpy
    x = H.Vector(np.zeros(N)); solvers.jacobi(A, b, x, 1, omega=1.0)
    assert np.allclose(host(x.values), host(b.values) / np.diag(_dense(N, rp, ci, va)), rtol=1e-15, atol=0)
