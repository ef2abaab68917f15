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


def _colouring_is_valid(rp, ci, color):
    """no stored entry (i, j), i != j, joins two rows of one colour - from either side"""
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    off = rows != ci
    return not np.any(color[rows[off]] == color[ci[off]])


@pytest.mark.parametrize("which", ["stencil", "stencil40", "lap5", "nonsymmetric"])
def test_symgs_matches_the_oracle_bit_for_bit(thsp, cuda, oracle, which):
    """thsp_symgs_f64 against oracle_symgs walking the same colours serially: same bits (rows of a colour do not touch each
    other, every row is added in stored order with unfused arithmetic).  The colouring itself is checked on the CPU."""
    from arm_spmv_b200 import host as H, solvers
    if which in ("stencil", "stencil40"):
        n = 14 if which == "stencil" else 40; N = n ** 3
        rp, ci, va = oracle.gen_stencil27_csr(n)
    elif which == "lap5":
        ri, cj, v = oracle.gen_lap5_coo(37); N = 37 * 37
        rp, ci, va, _ = oracle.coo2csr(N, N, ri, cj, v)
    else:   # random pattern, structurally NON-symmetric, with a dominant diagonal and a few empty-but-for-the-diagonal rows
        N = 3000
        ri, cj, v = oracle.gen_uniform_coo(N, N, 15000, 91)
        keep = ri != cj
        ri = np.concatenate([ri[keep], np.arange(N, dtype=np.int32)]); cj = np.concatenate([cj[keep], np.arange(N, dtype=np.int32)])
        v = np.concatenate([v[keep] - 0.5, np.full(N, 12.0)])
        rp, ci, va, _ = oracle.coo2csr(N, N, ri, cj, v)
    A = H.CSRMatrix(nrow=N, ncol=N, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    S = solvers.SymGS(A)
    cp, perm, color = S.coloring()
    assert S.ncolors <= 64 and cp[0] == 0 and cp[-1] == N and sorted(perm.tolist()) == list(range(N))
    assert _colouring_is_valid(rp, ci, color)
    for c in range(S.ncolors):
        assert np.all(color[perm[cp[c]:cp[c + 1]]] == c)
    if which in ("stencil", "stencil40"):
        assert S.ncolors >= 8        # a 27-point stencil needs 8
    # 64000 rows: the plan keeps the matrix permuted by colour and the colours go through the TMA stream kernel
    assert S.streams == (which == "stencil40")
    diag = host(S.diagonal.values)
    r = oracle.gen_vector(N, 5) - 0.5
    x0 = oracle.gen_vector(N, 6)
    x = H.Vector(x0)
    S.sweep(H.Vector(r), x)
    want = oracle.symgs(cp, perm, rp, ci, va, diag, r, x0)
    assert host(x.values).tobytes() == want.tobytes()
    # and it is a smoother: it brings A x closer to r than the sequential sweep's starting point
    res = lambda xx: np.linalg.norm(r - oracle.csr_spmv(N, N, rp, ci, va, xx, np.zeros(N)))
    assert res(want) < 0.5 * res(x0)
    seq = oracle.symgs_sequential(rp, ci, va, diag, r, x0)
    assert res(want) < 3.0 * res(seq) + 1e-12   # a reordering of the same method, not a worse one


@pytest.mark.parametrize("precond", ["none", "jacobi", "symgs"])
def test_cg_in_the_library_matches_the_oracle_iterates(thsp, cuda, oracle, precond):
    """thsp_cg_f64 (whole loop behind the C ABI) against oracle_cg: SpMV in the reference's order, the reference's vector
    forms, canonical dots - after a fixed number of iterations x has the same bits; run to convergence it solves the
    system, and the SymGS-preconditioned loop needs fewer iterations than the plain one."""
    from arm_spmv_b200 import host as H, solvers
    n = 12; N = n ** 3
    rp, ci, va = oracle.gen_stencil27_csr(n)
    A = H.CSRMatrix(nrow=N, ncol=N, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    assert A.plan_kernel()[0] in ("stream", "scalar")
    b = oracle.gen_vector(N, 3)
    M = solvers.SymGS(A) if precond == "symgs" else None
    d = solvers.csr_diagonal(A)
    cp, perm = (M.coloring()[:2] if M is not None else (None, None))
    kind = {"none": 0, "jacobi": 1, "symgs": 2}[precond]
    x = H.Vector(np.zeros(N))
    it, rel = solvers.pcg(A, H.Vector(b), x, tol=0.0, maxit=7, precond=precond, M=M, diagonal=d)
    want, wit, wrel = oracle.cg(rp, ci, va, host(d.values), b, np.zeros(N), 7, 0.0, kind, cp, perm)
    assert it == wit == 7
    assert host(x.values).tobytes() == want.tobytes()
    assert rel == wrel
    x = H.Vector(np.zeros(N))
    it, rel = solvers.pcg(A, H.Vector(b), x, tol=1e-11, maxit=500, precond=precond, M=M, diagonal=d)
    assert rel <= 1e-11
    resid = b - oracle.csr_spmv(N, N, rp, ci, va, host(x.values), np.zeros(N))
    assert np.linalg.norm(resid) <= 1e-10 * np.linalg.norm(b)
    if precond == "symgs":
        it0, _ = solvers.pcg(A, H.Vector(b), H.Vector(np.zeros(N)), tol=1e-11, maxit=500)
        assert it < it0
        # the streamed sweep inside the loop: 40^3, five iterations, same bits as the serial twin
        n2 = 40; N2 = n2 ** 3
        rp2, ci2, va2 = oracle.gen_stencil27_csr(n2)
        A2 = H.CSRMatrix(nrow=N2, ncol=N2, row_ptr=dev(rp2), col_ind=dev(ci2), values=dev(va2))
        M2 = solvers.SymGS(A2)
        assert M2.streams
        cp2, perm2, _ = M2.coloring()
        b2 = oracle.gen_vector(N2, 8)
        x2 = H.Vector(np.zeros(N2))
        solvers.pcg(A2, H.Vector(b2), x2, tol=0.0, maxit=5, precond="symgs", M=M2)
        want2, _, _ = oracle.cg(rp2, ci2, va2, host(M2.diagonal.values), b2, np.zeros(N2), 5, 0.0, 2, cp2, perm2)
        assert host(x2.values).tobytes() == want2.tobytes()


def test_dot_canonical(thsp, cuda, oracle):
    import ctypes as C
    import torch
    from arm_spmv_b200.lib import check, current_stream, ptr
    lib = thsp.load()
    for n in (1, 33, 100_003):
        a, b = oracle.gen_vector(n, 1) - 0.5, oracle.gen_vector(n, 2)
        da, db = dev(a), dev(b)
        tiles = torch.empty((n + 31) // 32, dtype=torch.float64, device="cuda")
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        check(lib.thsp_dot_canonical_dev_f64(C.c_int64(n), ptr(da), ptr(db), ptr(tiles), ptr(out), current_stream()))
        assert float(out.item()) == oracle.dot_canonical(a, b)
