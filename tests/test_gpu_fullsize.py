"""-m gpu, BASELINE.json's FULL sizes: the CPU oracle cannot finish these in seconds, so the kernels are checked through
properties that do not depend on the size - known answers (row sums of the stencil), kernels whose summation order is
the reference's agreeing bit for bit with each other, linearity, and the conversions against an independent stable
sort (torch.sort(stable=True), checker only) - the reference's counting sort IS a stable sort by row / column
(src/matrix.cpp:125-144)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL64 = 1e-12


def _csr_kernel(H, kernel, lanes, A, x, accumulate=False):
    y = torch.zeros(A.nrow, dtype=torch.float64, device="cuda")
    H.csr_spmv_kernel(kernel, lanes, A, x, y, accumulate)
    return y


def _row_scale(H, A, x):
    """sum_j |a_ij x_j| per row, by the in-order scalar kernel on |A|, |x| (the error denominator)."""
    B = H.CSRMatrix(nrow=A.nrow, ncol=A.ncol, row_ptr=A.row_ptr, col_ind=A.col_ind, values=A.values.abs())
    return _csr_kernel(H, 1, 1, B, x.abs())


def _max_row_err(y, ref, scale):
    den = scale.clone()
    err = (y - ref).abs()
    assert bool((err[den == 0] == 0).all()), "a row with no contributions changed"
    m = den > 0
    return float((err[m] / den[m]).max().item()) if bool(m.any()) else 0.0


def test_stencil_256_known_answer_and_bitwise_agreement(thsp, cuda):
    """configs[1]: 27-point stencil 256^3.  A * ones = 27 - (row length) exactly (26 on the diagonal, -1 elsewhere);
    CSR stream, CSR scalar, ELL and DIA all add a row's products in stored order from 0, so with y0 = 0 their results
    are the same bits on a random x as well."""
    from arm_spmv_b200 import host as H
    n = 256
    N = n ** 3
    A = H.stencil27_csr(n)
    assert A.nnz == (3 * n - 2) ** 3 and int(A.row_ptr[-1]) == A.nnz
    lens = (A.row_ptr[1:] - A.row_ptr[:-1]).to(torch.float64)
    ones = torch.ones(N, dtype=torch.float64, device="cuda")
    want = 27.0 - lens
    x = H.gen_vector(N, 5).values
    outs = {}
    for name, (k, l) in {"stream": (3, 1), "scalar": (1, 1)}.items():
        assert torch.equal(_csr_kernel(H, k, l, A, ones), want), name
        outs[name] = _csr_kernel(H, k, l, A, x)
    assert torch.equal(outs["stream"], outs["scalar"])
    for k, l in [(2, 8), (4, 1)]:   # vector / merge-path: other orders, exact on integers, tolerance on x
        assert torch.equal(_csr_kernel(H, k, l, A, ones), want)
    scale = _row_scale(H, A, x)
    assert _max_row_err(_csr_kernel(H, 4, 1, A, x), outs["stream"], scale) <= TOL64
    assert _max_row_err(_csr_kernel(H, 2, 8, A, x), outs["stream"], scale) <= TOL64
    del scale
    E = H.stencil27_ell(n)
    ye = H.Vector(N); ye.Fill(0.0)
    H.ELLMatrixMatVector(E, H.Vector(x), ye)
    assert torch.equal(ye.values, outs["stream"])
    del E, ye
    D = H.DIAMatrix(A)
    assert D.ndiags == 27
    yd = H.Vector(N); yd.Fill(0.0)
    H.DIAMatrixMatVector(D, H.Vector(x), yd)
    assert torch.equal(yd.values, outs["stream"])
    # accumulate: y0 + A x with the sum formed first (src/mat_vec.cpp:58-64)
    y0 = H.gen_vector(N, 6).values
    ya = y0.clone()
    H.csr_spmv_kernel(3, 1, A, x, ya, True)
    assert torch.equal(ya, y0 + outs["stream"])


def _check_compressed(H, nb, key, oth, val, ptr, out_oth, out_val):
    """(ptr, out_oth, out_val) must be the stable sort of the entries by key."""
    order = torch.sort(key.to(torch.int64), stable=True).indices
    assert torch.equal(out_oth, oth[order]) and torch.equal(out_val, val[order])
    counts = torch.bincount(key.to(torch.int64), minlength=nb)
    want_ptr = torch.zeros(nb + 1, dtype=torch.int64, device="cuda")
    want_ptr[1:] = torch.cumsum(counts, 0)
    assert torch.equal(ptr.to(torch.int64), want_ptr)
    return order


@pytest.mark.parametrize("which", ["uniform", "rmat"])
def test_conversions_full_size_are_the_stable_sort(thsp, cuda, which):
    """configs[3] (uniform 8M x 8M, 128 M entries) and configs[2] (R-MAT scale 24, 268 M entries, duplicates kept):
    COO -> CSR / CSC index and value arrays bit for bit, ELL slab for the uniform matrix."""
    from arm_spmv_b200 import host as H
    A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43) if which == "uniform" else H.rmat_coo(24, 16 << 24, 42)
    # values = original position: the output then shows the permutation itself
    A.values.copy_(torch.arange(A.nnz, dtype=torch.float64, device="cuda"))
    B = H.CSRMatrix(A)
    order = _check_compressed(H, A.nrow, A.row_ind, A.col_ind, A.values, B.row_ptr, B.col_ind, B.values)
    assert torch.equal(B.values.to(torch.int64), order)
    if which == "uniform":
        D = H.ELLMatrix(A)
        K = D.nonzeros_in_row
        lens = B.row_ptr[1:] - B.row_ptr[:-1]
        assert K == int(lens.max())
        slab_c = D.col_ind.view(K, A.nrow)
        slab_v = D.values.view(K, A.nrow)
        for k in (0, 1, K // 2, K - 1):   # slot k of every row: entry row_ptr[r] + k, or padding (column 0, 0.0)
            has = lens > k
            src = (B.row_ptr[:-1].to(torch.int64) + k)[has]
            assert torch.equal(slab_c[k][has], B.col_ind[src]) and torch.equal(slab_v[k][has], B.values[src])
            assert bool((slab_c[k][~has] == 0).all()) and bool((slab_v[k][~has] == 0).all())
        del D, slab_c, slab_v
    del B, order
    Cc = H.CSCMatrix(A)
    _check_compressed(H, A.ncol, A.col_ind, A.row_ind, A.values, Cc.col_ptr, Cc.row_ind, Cc.values)


def test_stencil_256_coo_to_csc_is_transposed_not_sorted(thsp, cuda):
    """configs[1] as a COO matrix (449 M entries, row by row): COO -> CSC goes through transpose_entries (convert.cu:
    per-column cursors, entry numbers in the slots, per-column sort, gather), not through the radix sort, and its arrays are
    the stable sort by column bit for bit (values = original position, so the output shows the permutation itself).
    src/matrix.cpp:295-325."""
    from arm_spmv_b200 import host as H
    A = H.stencil27_coo(256)
    A.values.copy_(torch.arange(A.nnz, dtype=torch.float64, device="cuda"))
    Cc = H.CSCMatrix(A)
    assert thsp.load().thsp_coo_last_path() == 2
    order = _check_compressed(H, A.ncol, A.col_ind, A.row_ind, A.values, Cc.col_ptr, Cc.row_ind, Cc.values)
    assert torch.equal(Cc.values.to(torch.int64), order)
    del order, Cc
    B = H.CSRMatrix(A)   # the keys are in order already: copied through
    assert thsp.load().thsp_coo_last_path() == 0
    assert torch.equal(B.col_ind, A.col_ind) and torch.equal(B.values, A.values)


@pytest.mark.parametrize("which", ["uniform", "rmat"])
def test_spmv_full_size_formats_agree(thsp, cuda, which):
    """Every CSR kernel, COO and CSC on the full-size irregular matrices against the in-order scalar CSR kernel
    (the reference's summation order), per-row error <= 1e-12; fp32 merge-path <= 1e-5; linearity of the plan kernel."""
    from arm_spmv_b200 import host as H
    A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43) if which == "uniform" else H.rmat_coo(24, 16 << 24, 42)
    B = H.CSRMatrix(A)
    x = H.gen_vector(A.ncol, 3).values - 0.5
    ref = _csr_kernel(H, 1, 1, B, x)
    scale = _row_scale(H, B, x)
    kernels = [(4, 1), (2, 32)] + ([(3, 1), (2, 8)] if which == "uniform" else [])
    for k, l in kernels:
        assert _max_row_err(_csr_kernel(H, k, l, B, x), ref, scale) <= TOL64, (k, l)
    yp = H.Vector(B.nrow); yp.Fill(0.0)
    H.CSRMatrixMatVector(B, H.Vector(x), yp)                        # the plan's choice
    assert _max_row_err(yp.values, ref, scale) <= TOL64
    z = H.gen_vector(A.ncol, 9).values
    lin = _csr_kernel(H, 0, 0, B, 2.0 * x + z)                      # A (2x + z) = 2 A x + A z
    rz = _csr_kernel(H, 0, 0, B, z)
    scale2 = _row_scale(H, B, 2.0 * x.abs() + z.abs())
    assert _max_row_err(lin, 2.0 * _csr_kernel(H, 0, 0, B, x) + rz, scale2) <= 4 * TOL64
    del lin, rz, scale2, z
    yc = H.Vector(A.nrow); yc.Fill(0.0)
    H.COOMatirxMatVector(A, H.Vector(x), yc)
    assert _max_row_err(yc.values, ref, scale) <= TOL64
    Cc = H.CSCMatrix(A)
    yc.Fill(0.0)
    H.CSCMatrixMatVector(Cc, H.Vector(x), yc)
    assert _max_row_err(yc.values, ref, scale) <= TOL64
    del Cc, yc
    B32 = H.CSRMatrix(nrow=B.nrow, ncol=B.ncol, row_ptr=B.row_ptr, col_ind=B.col_ind, values=B.values.to(torch.float32))
    y32 = torch.zeros(B.nrow, dtype=torch.float32, device="cuda")
    H.csr_spmv_kernel(4, 1, B32, x.to(torch.float32), y32, False)
    ref32 = _csr_kernel(H, 1, 1, H.CSRMatrix(nrow=B.nrow, ncol=B.ncol, row_ptr=B.row_ptr, col_ind=B.col_ind,
                                             values=B.values.to(torch.float32).to(torch.float64)), x.to(torch.float32).to(torch.float64))
    assert _max_row_err(y32.to(torch.float64), ref32, scale) <= 1e-5


# ------------------------------------------------------------------------------------------------------------------
# The same full sizes against the REFERENCE ITSELF (oracle/_ref/libref.so = the unmodified sources compiled by
# oracle/Makefile; it travels to the GPU box prebuilt), the C restatement when that library is absent.
def _checker():
    import pyoracle
    pyoracle.build()
    try:
        r = pyoracle.Ref()
        r.set_threads(max(1, min(16, len(__import__("os").sched_getaffinity(0)))))
        return r, "reference"
    except (FileNotFoundError, OSError):
        return pyoracle.Oracle(), "port"


def _np(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def test_conversions_config3_equal_the_reference_constructors(thsp, cuda):
    """configs[3], 128 M unsorted entries: CSRMatrix(COO), CSCMatrix(COO), ELLMatrix(COO) of the reference
    (src/matrix.cpp:115-154, 295-325, 450-500; serial counting sorts, a few seconds each on the host) against the
    GPU conversions - every index, value and diagonal array compared byte for byte."""
    from arm_spmv_b200 import host as H
    from gpu_util import assert_bits
    chk, kind = _checker()
    nrow = ncol = 1 << 23
    A = H.uniform_coo(nrow, ncol, 1 << 27, 43)
    ri, ci, va = _np(A.row_ind), _np(A.col_ind), _np(A.values)
    B = H.CSRMatrix(A)
    rp, co, vo, dg = chk.coo2csr(nrow, ncol, ri, ci, va)
    assert_bits(_np(B.row_ptr), rp, f"row_ptr vs {kind}")
    assert_bits(_np(B.col_ind), co, f"CSR col_ind vs {kind}")
    assert_bits(_np(B.values), vo, f"CSR values vs {kind}")
    assert B.ndiag == len(dg)
    assert_bits(_np(B.diagonal[:B.ndiag]), dg, f"CSR diagonal vs {kind}")
    del B, rp, co, vo
    Cc = H.CSCMatrix(A)
    cp, ro, vo = chk.coo2csc(nrow, ncol, ri, ci, va)
    assert_bits(_np(Cc.col_ptr), cp, f"col_ptr vs {kind}")
    assert_bits(_np(Cc.row_ind), ro, f"CSC row_ind vs {kind}")
    assert_bits(_np(Cc.values), vo, f"CSC values vs {kind}")
    del Cc, cp, ro, vo
    D = H.ELLMatrix(A)
    k, eco, eva, edg = chk.coo2ell(nrow, ncol, ri, ci, va)
    assert D.nonzeros_in_row == k
    assert_bits(_np(D.col_ind), eco, f"ELL col_ind vs {kind}")
    assert_bits(_np(D.values), eva, f"ELL values vs {kind}")
    assert_bits(_np(D.diagonal[:D.ndiag]), edg, f"ELL diagonal vs {kind}")


def test_stencil_256_spmv_equals_the_reference_bit_for_bit(thsp, cuda, oracle):
    """configs[1]: y += A x of the reference's own CSRMatrixMatVector / ELLMatrixMatVector / DIAMatrixMatVector
    (src/mat_vec.cpp:44-67, 97-121, 123-146) on the 256^3 stencil against the GPU kernels, same bits.  The matrix
    arrays the GPU generator makes are first checked against the oracle's generator."""
    from arm_spmv_b200 import host as H
    from gpu_util import assert_bits
    chk, kind = _checker()
    n = 256
    N = n ** 3
    rp, ci, va = oracle.gen_stencil27_csr(n)
    xh = oracle.gen_vector(N, 11)
    A = H.stencil27_csr(n)
    x = H.gen_vector(N, 11)
    assert_bits(_np(A.row_ptr), rp, "stencil row_ptr")
    assert_bits(_np(A.col_ind), ci, "stencil col_ind")
    assert_bits(_np(A.values), va, "stencil values")
    assert_bits(_np(x.values), xh, "x")
    y0 = oracle.gen_vector(N, 12)
    want = chk.csr_spmv(N, N, rp, ci, va, xh, y0)
    for kernel in (3, 1):   # stream (the plan's choice) and scalar: the reference's order
        y = H.Vector(y0)
        H.csr_spmv_kernel(kernel, 1, A, x.values, y.values, True)
        assert_bits(_np(y.values), want, f"CSR kernel {kernel} vs {kind}")
    del rp, ci, va
    D = H.DIAMatrix(A)
    yd = H.Vector(y0)
    H.DIAMatrixMatVector(D, x, yd)
    wantd = chk.dia_spmv(N, N, _np(D.offsets), _np(D.values), xh, y0)
    assert_bits(_np(yd.values), wantd, f"DIA vs {kind}")
    del D, A, wantd
    E = H.stencil27_ell(n)
    ye = H.Vector(y0)
    H.ELLMatrixMatVector(E, x, ye)
    wante = chk.ell_spmv(N, N, 27, _np(E.col_ind), _np(E.values), xh, y0)
    assert_bits(_np(ye.values), wante, f"ELL vs {kind}")
