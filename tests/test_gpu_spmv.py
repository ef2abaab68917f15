"""-m gpu parity: every SpMV kernel (through the C ABI) against the CPU oracle and the golden
fixtures generated from the unmodified reference.

Tolerances (north star): fp64 per-row error <= 1e-12, fp32 <= 1e-5, where the per-row error is
|y - y_ref| / (sum_j |a_ij x_j| + |y0|).  Kernels that keep the reference's
summation order (CSR scalar/stream, ELL, DIA) must be BIT-IDENTICAL to it."""
import numpy as np
import pytest
import torch

import cases as C
from conftest import load_golden
from gpu_util import assert_bits, dev, host, max_row_error, row_scale_coo, row_scale_csr

pytestmark = pytest.mark.gpu
TOL64 = 1e-12
TOL32 = 1e-5
NAMES = list(C.cases().keys())
SCALAR, VECTOR, STREAM, MERGE = 1, 2, 3, 4
CSR_KERNELS = [(SCALAR, 1), (VECTOR, 2), (VECTOR, 4), (VECTOR, 8), (VECTOR, 16), (VECTOR, 32), (STREAM, 1), (MERGE, 1)]
EXACT = {(SCALAR, 1), (STREAM, 1)}


def run_csr(thsp, kernel, lanes, nrow, ncol, rp, ci, va, x, y0, accumulate=True, dtype=torch.float64):
    from arm_spmv_b200 import host as H
    A = H.CSRMatrix(nrow=nrow, ncol=ncol, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va, dtype))
    xd, yd = dev(x, dtype), dev(y0, dtype)
    H.csr_spmv_kernel(kernel, lanes, A, xd, yd, accumulate)
    return host(yd)


def synthetic(oracle, which):
    if which == "stencil12":
        rp, ci, va = oracle.gen_stencil27_csr(12)
        return 1728, 1728, rp, ci, va
    if which == "lap5_40":
        ri, cj, v = oracle.gen_lap5_coo(40)
        rp, ci, va, _ = oracle.coo2csr(1600, 1600, ri, cj, v)
        return 1600, 1600, rp, ci, va
    if which == "rmat12":
        ri, cj, v = oracle.gen_rmat_coo(12, 60000, 42)
        rp, ci, va, _ = oracle.coo2csr(4096, 4096, ri, cj, v)
        return 4096, 4096, rp, ci, va
    if which == "uniform":
        ri, cj, v = oracle.gen_uniform_coo(3000, 2500, 20011, 43)
        rp, ci, va, _ = oracle.coo2csr(3000, 2500, ri, cj, v)
        return 3000, 2500, rp, ci, va
    if which == "hub":  # one row holding 70% of a 200k-entry matrix: runs far longer than a merge tile
        rs = np.random.RandomState(5)
        nrow = ncol = 5000
        ri = rs.randint(0, nrow, 200000).astype(np.int32); ri[:140000] = 1234
        cj = rs.randint(0, ncol, 200000).astype(np.int32); v = rs.uniform(-1, 1, 200000)
        rp, ci, va, _ = oracle.coo2csr(nrow, ncol, ri, cj, v)
        return nrow, ncol, rp, ci, va
    raise KeyError(which)


@pytest.mark.parametrize("kernel,lanes", CSR_KERNELS)
@pytest.mark.parametrize("name", NAMES)
def test_csr_golden(thsp, cuda, name, kernel, lanes):
    g = load_golden(name)
    nrow, ncol = int(g["nrow"]), int(g["ncol"])
    rp, ci, va, x, y0 = g["csr_row_ptr"], g["csr_col_ind"], g["csr_values"], g["x"], g["y0"]
    y = run_csr(thsp, kernel, lanes, nrow, ncol, rp, ci, va, x, y0)
    if (kernel, lanes) in EXACT:
        assert_bits(y, g["y_csr"], f"csr {name} k{kernel}")
    else:
        assert max_row_error(y, g["y_csr"], row_scale_csr(nrow, rp, ci, va, x), y0) <= TOL64


@pytest.mark.parametrize("kernel,lanes", CSR_KERNELS)
@pytest.mark.parametrize("which", ["stencil12", "lap5_40", "rmat12", "uniform", "hub"])
@pytest.mark.parametrize("accumulate", [True, False])
def test_csr_synthetic(thsp, cuda, oracle, which, kernel, lanes, accumulate):
    nrow, ncol, rp, ci, va = synthetic(oracle, which)
    x = oracle.gen_vector(ncol, 7)
    y0 = oracle.gen_vector(nrow, 8) - 0.5
    ref = oracle.csr_spmv(nrow, ncol, rp, ci, va, x, y0 if accumulate else np.zeros(nrow))
    y = run_csr(thsp, kernel, lanes, nrow, ncol, rp, ci, va, x, y0, accumulate)
    if (kernel, lanes) in EXACT:
        assert_bits(y, ref, f"csr {which} k{kernel} acc{accumulate}")
    else:
        assert max_row_error(y, ref, row_scale_csr(nrow, rp, ci, va, x), y0 if accumulate else None) <= TOL64


def test_csr_stream_ragged_alignment(thsp, cuda, oracle):
    """Stream kernel: tiles that start at every residue mod 4, nnz not a multiple of 4, rows longer
    than one shared-memory stage, and sub-array views whose base is only 16 B aligned."""
    rs = np.random.RandomState(3)
    for trial in range(6):
        nrow = int(rs.randint(1, 400)); ncol = int(rs.randint(1, 300))
        lens = rs.randint(0, 9, nrow)
        if trial % 2:
            lens[rs.randint(0, nrow)] = 5000  # several chunks for one tile
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(rp[-1])
        ci = rs.randint(0, ncol, nnz).astype(np.int32); va = rs.uniform(-1, 1, nnz)
        x = rs.uniform(0, 1, ncol); y0 = rs.uniform(-1, 1, nrow)
        y = run_csr(thsp, STREAM, 1, nrow, ncol, rp, ci, va, x, y0)
        assert_bits(y, oracle.csr_spmv(nrow, ncol, rp, ci, va, x, y0), f"ragged trial {trial}")


def _merge_edge_lens(which, rs):
    R = 496  # merge steps per warp-run (kMbRun = 31 lane blocks of 16 entries)
    if which == "lane_aligned":      # every row ends exactly at a lane boundary
        return np.full(3000, 16)
    if which == "run_aligned":       # row + its end = exactly one run
        return np.full(200, R - 1)
    if which == "run_aligned_512":   # the same for the shared-memory variant (512 steps per run)
        return np.full(200, 511)
    if which == "empty_stretches":   # thousands of empty rows around one long row, then short rows
        return np.concatenate([np.zeros(5000, int), [3000], np.zeros(7000, int), rs.randint(0, 4, 2000), np.zeros(1500, int)])
    if which == "single_entry":
        return np.array([0, 0, 1, 0])
    if which == "mixed":
        return rs.choice([0, 0, 0, 1, 2, 15, 16, 17, 40, 495, 496, 497, 511, 512, 513, 700], 4000)
    if which == "one_long_row":
        return np.array([100000])
    if which == "all_empty_but_last":
        return np.concatenate([np.zeros(20000, int), [5]])
    raise KeyError(which)


@pytest.mark.parametrize("which", ["lane_aligned", "run_aligned", "run_aligned_512", "empty_stretches", "single_entry", "mixed", "one_long_row",
                                   "all_empty_but_last"])
@pytest.mark.parametrize("accumulate", [True, False])
def test_csr_merge_path_edges(thsp, cuda, oracle, which, accumulate):
    """Merge-path kernel: rows ending on lane / run boundaries, long stretches of empty rows, rows spanning
    many runs; through the stateless entry point and through the plan (cached run table)."""
    from arm_spmv_b200 import host as H
    rs = np.random.RandomState(11)
    lens = np.asarray(_merge_edge_lens(which, rs), dtype=np.int64)
    nrow = len(lens); ncol = 777
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nnz = int(rp[-1])
    ci = rs.randint(0, ncol, nnz).astype(np.int32); va = rs.uniform(-1, 1, nnz)
    x = rs.uniform(0, 1, ncol); y0 = rs.uniform(-1, 1, nrow)
    ref = oracle.csr_spmv(nrow, ncol, rp, ci, va, x, y0 if accumulate else np.zeros(nrow))
    scale = row_scale_csr(nrow, rp, ci, va, x)
    y = run_csr(thsp, MERGE, 1, nrow, ncol, rp, ci, va, x, y0, accumulate)
    assert max_row_error(y, ref, scale, y0 if accumulate else None) <= TOL64
    A = H.CSRMatrix(nrow=nrow, ncol=ncol, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    thsp.lib.check(thsp.load().thsp_csr_plan_set_kernel(A.plan(), MERGE, 1))
    for _ in range(2):   # second call reuses the cached run table
        xv, yv = H.Vector(x), H.Vector(y0.copy())
        H.CSRMatrixMatVector(A, xv, yv, accumulate)
        assert max_row_error(host(yv.values), ref, scale, y0 if accumulate else None) <= TOL64


def test_csr_plan_picks_kernel_from_histogram(thsp, cuda, oracle):
    from arm_spmv_b200 import host as H
    nrow, ncol, rp, ci, va = synthetic(oracle, "stencil12")
    A = H.CSRMatrix(nrow=nrow, ncol=ncol, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    assert A.plan_kernel()[0] == "stream"
    nrow, ncol, rp, ci, va = synthetic(oracle, "hub")
    B = H.CSRMatrix(nrow=nrow, ncol=ncol, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    assert B.plan_kernel()[0] == "merge"
    # plan path result == oracle
    x = H.Vector(oracle.gen_vector(ncol, 1)); y = H.Vector(np.zeros(nrow))
    H.CSRMatrixMatVector(B, x, y)
    ref = oracle.csr_spmv(nrow, ncol, rp, ci, va, host(x.values), np.zeros(nrow))
    assert max_row_error(host(y.values), ref, row_scale_csr(nrow, rp, ci, va, host(x.values))) <= TOL64
    # histogram bins: 2^(b-1) <= len < 2^b
    import ctypes
    hist = (ctypes.c_int64 * 32)(); mx = ctypes.c_int()
    thsp.lib.check(thsp.load().thsp_csr_plan_histogram(B.plan(), hist, ctypes.byref(mx)))
    lens = np.diff(rp)
    assert mx.value == lens.max() and sum(hist) == nrow and hist[0] == int((lens == 0).sum())


@pytest.mark.parametrize("kernel,lanes", [(SCALAR, 1), (VECTOR, 8), (STREAM, 1), (MERGE, 1)])
def test_csr_fp32(thsp, cuda, oracle, kernel, lanes):
    nrow, ncol, rp, ci, va = synthetic(oracle, "stencil12")
    x = oracle.gen_vector(ncol, 7).astype(np.float32); va32 = va.astype(np.float32)
    y0 = np.zeros(nrow, np.float32)
    y = run_csr(thsp, kernel, lanes, nrow, ncol, rp, ci, va32, x, y0, True, torch.float32)
    ref32 = oracle.csr_spmv_f32(nrow, rp, ci, va32, x, y0)
    if (kernel, lanes) in EXACT:
        assert_bits(y, ref32, "fp32 in-order")
    # fp32 oracle per SURVEY.md 8(c): the fp64 reference sum of the fp32-rounded inputs
    ref64 = oracle.csr_spmv(nrow, ncol, rp, ci, va32.astype(np.float64), x.astype(np.float64), np.zeros(nrow))
    assert max_row_error(y.astype(np.float64), ref64, row_scale_csr(nrow, rp, ci, va32.astype(np.float64), x.astype(np.float64))) <= TOL32


@pytest.mark.parametrize("name", NAMES)
def test_ell_coo_csc_dia_golden(thsp, cuda, name):
    from arm_spmv_b200 import host as H
    g = load_golden(name)
    nrow, ncol = int(g["nrow"]), int(g["ncol"])
    x, y0 = g["x"], g["y0"]
    X = H.Vector(x)
    scale = row_scale_coo(nrow, g["ri"], g["ci"], g["va"], x)
    # ELL: reference order kept -> bit-identical
    E = H.ELLMatrix(nrow=nrow, ncol=ncol, nnz=len(g["va"]), nonzeros_in_row=int(g["ell_width"]), col_ind=g["ell_col_ind"],
                    values=g["ell_values"])
    Y = H.Vector(y0); H.ELLMatrixMatVector(E, X, Y)
    assert_bits(host(Y.values), g["y_ell"], f"ell {name}")
    # COO / CSC: atomics -> tolerance
    A = H.COOMatrix(nrow, ncol, g["ri"], g["ci"], g["va"])
    Y = H.Vector(y0); H.COOMatirxMatVector(A, X, Y)
    assert max_row_error(host(Y.values), g["y_coo"], scale, y0) <= TOL64
    Cm = H.CSCMatrix(nrow=nrow, ncol=ncol, col_ptr=g["csc_col_ptr"], row_ind=g["csc_row_ind"], values=g["csc_values"])
    Y = H.Vector(y0); H.CSCMatrixMatVector(Cm, X, Y)
    assert max_row_error(host(Y.values), g["y_csc"], scale, y0) <= TOL64
    if "y_dia" in g:
        D = H.DIAMatrix(nrow=nrow, ncol=ncol, offsets=g["dia_offsets"], values=g["dia_values"])
        Y = H.Vector(y0); H.DIAMatrixMatVector(D, X, Y)
        assert_bits(host(Y.values), g["y_dia"], f"dia {name}")


@pytest.mark.parametrize("which", ["stencil12", "lap5_40", "rmat12", "uniform"])
def test_formats_synthetic(thsp, cuda, oracle, which):
    """All five formats built on the GPU from one COO agree with the oracle's SpMV."""
    from arm_spmv_b200 import host as H
    nrow, ncol, rp, ci, va = synthetic(oracle, which)
    ri = np.repeat(np.arange(nrow, dtype=np.int32), np.diff(rp))
    rs = np.random.RandomState(1); perm = rs.permutation(len(va))      # unsorted COO
    ri, cj, v = ri[perm], ci[perm], va[perm]
    x = oracle.gen_vector(ncol, 3); y0 = oracle.gen_vector(nrow, 4)
    scale = row_scale_coo(nrow, ri, cj, v, x)
    A = H.COOMatrix(nrow, ncol, ri, cj, v); X = H.Vector(x)
    ref = oracle.coo_spmv(nrow, ncol, ri, cj, v, x, y0)
    Y = H.Vector(y0); H.COOMatirxMatVector(A, X, Y)
    assert max_row_error(host(Y.values), ref, scale, y0) <= TOL64
    B = H.CSRMatrix(A); Y = H.Vector(y0); H.CSRMatrixMatVector(B, X, Y)
    rp2, ci2, va2, _ = oracle.coo2csr(nrow, ncol, ri, cj, v)
    refc = oracle.csr_spmv(nrow, ncol, rp2, ci2, va2, x, y0)
    assert max_row_error(host(Y.values), refc, scale, y0) <= TOL64
    Cc = H.CSCMatrix(A); Y = H.Vector(y0); H.CSCMatrixMatVector(Cc, X, Y)
    assert max_row_error(host(Y.values), ref, scale, y0) <= TOL64
    if which != "rmat12":   # ELL width of a power-law matrix is its longest row: keep the slab small
        D = H.ELLMatrix(A); Y = H.Vector(y0); H.ELLMatrixMatVector(D, X, Y)
        k, eco, eva, _ = oracle.coo2ell(nrow, ncol, ri, cj, v)
        assert_bits(host(Y.values), oracle.ell_spmv(nrow, ncol, k, eco, eva, x, y0), f"ell {which}")
    if which in ("stencil12", "lap5_40"):
        E = H.DIAMatrix(B); Y = H.Vector(y0); H.DIAMatrixMatVector(E, X, Y)
        off, dv = oracle.csr2dia(nrow, ncol, rp2, ci2, va2)
        assert_bits(host(Y.values), oracle.dia_spmv(nrow, ncol, off, dv, x, y0), f"dia {which}")


@pytest.mark.parametrize("nnz", [0, 1, 15, 16, 17, 2047, 2048, 2049, 100_003])
@pytest.mark.parametrize("path", [0, 1])
def test_coo_paths_match_oracle(thsp, cuda, oracle, nnz, path):
    """Both COO paths (fused; products then scatter) against COOMatirxMatVector (src/mat_vec.cpp:18-42) around the
    kernels' block sizes (16 entries per lane, 2048 per CTA of the product kernel), y += semantics, unsorted entries."""
    lib = thsp.load()
    nrow, ncol = 5000, 4000
    ri, cj, v = oracle.gen_uniform_coo(nrow, ncol, max(nnz, 1), 43)
    ri, cj, v = ri[:nnz], cj[:nnz], v[:nnz]
    x = oracle.gen_vector(ncol, 5); y0 = oracle.gen_vector(nrow, 6) - 0.5
    y = dev(y0.copy()); d_ri, d_cj, d_v, d_x = dev(ri), dev(cj), dev(v), dev(x)
    thsp.lib.check(lib.thsp_coo_spmv_path_f64(path, nrow, ncol, nnz, thsp.lib.ptr(d_ri), thsp.lib.ptr(d_cj), thsp.lib.ptr(d_v),
                                              thsp.lib.ptr(d_x), thsp.lib.ptr(y), thsp.lib.current_stream()))
    torch.cuda.synchronize()
    ref = oracle.coo_spmv(nrow, ncol, ri, cj, v, x, y0)
    assert max_row_error(host(y), ref, row_scale_coo(nrow, ri, cj, v, x), y0) <= TOL64


def test_coo_sorted_uses_carry(thsp, cuda, oracle):
    """Row-sorted COO: same answer as CSR within tolerance (segments chained across steps)."""
    from arm_spmv_b200 import host as H
    rp, ci, va = oracle.gen_stencil27_csr(10)
    nrow = 1000
    ri = np.repeat(np.arange(nrow, dtype=np.int32), np.diff(rp))
    x = oracle.gen_vector(nrow, 9)
    Y = H.Vector(np.zeros(nrow)); H.COOMatirxMatVector(H.COOMatrix(nrow, nrow, ri, ci, va), H.Vector(x), Y)
    ref = oracle.csr_spmv(nrow, nrow, rp, ci, va, x, np.zeros(nrow))
    assert max_row_error(host(Y.values), ref, row_scale_csr(nrow, rp, ci, va, x)) <= TOL64


@pytest.mark.parametrize("which,accumulate", [("stencil128", False), ("stencil128", True), ("uniform_big", False), ("tiny", True)])
def test_csr_host_buffer_path(thsp, cuda, oracle, which, accumulate):
    """thsp_csr_plan_spmv_host_f64: pinned x in, y out, row chunks pipelined over three streams.
    Must equal the device path bit for bit (same kernels, same per-row order)."""
    import ctypes
    from arm_spmv_b200 import host as H
    if which == "stencil128":      # 2.1 M rows -> 2 chunks, banded footprints
        rp, ci, va = oracle.gen_stencil27_csr(128); nrow = ncol = 128 ** 3
    elif which == "uniform_big":   # 3.2 M rows -> 3 chunks, every chunk needs all of x
        nrow = ncol = 3_200_000
        ri, cj, v = oracle.gen_uniform_coo(nrow, ncol, 12_000_000, 43)
        rp, ci, va, _ = oracle.coo2csr(nrow, ncol, ri, cj, v)
    else:
        rp, ci, va = oracle.gen_stencil27_csr(7); nrow = ncol = 343
    A = H.CSRMatrix(nrow=nrow, ncol=ncol, row_ptr=dev(rp), col_ind=dev(ci), values=dev(va))
    x = oracle.gen_vector(ncol, 21); y0 = oracle.gen_vector(nrow, 22) - 0.5
    xh = torch.from_numpy(x).pin_memory(); yh = torch.from_numpy(y0.copy()).pin_memory()
    xd = torch.full((ncol,), float("nan"), dtype=torch.float64, device="cuda")
    yd = torch.full((nrow,), float("nan"), dtype=torch.float64, device="cuda")
    lib = thsp.load()
    yh_first = yh
    yh_other = torch.from_numpy(y0.copy()).pin_memory()
    y_pageable = y0.copy()
    # call 1 runs eagerly and builds the pipeline, call 2 captures it as a CUDA graph, calls 3-4 replay the graph,
    # call 5 brings another pinned y (graph dropped, eager again), call 6 a pageable y (never captured); before calls 7-9
    # the plan's launch shape changes, so the chunks and the graph are rebuilt
    launches0 = thsp.lib.launch_count()
    for it in range(9):
        if it == 6:
            thsp.lib.check(lib.thsp_csr_plan_set_stream_config(A.plan(), 0, 0, 0, 100))
        yh = yh_first if (it < 4 or it > 5) else (yh_other if it == 4 else torch.from_numpy(y_pageable))
        yh.copy_(torch.from_numpy(y0))
        thsp.lib.check(lib.thsp_csr_plan_spmv_host_f64(A.plan(), ctypes.c_void_p(xh.data_ptr()), ctypes.c_void_p(yh.data_ptr()),
                                                       thsp.lib.ptr(xd), thsp.lib.ptr(yd), 1 if accumulate else 0, thsp.lib.current_stream()))
        ref = oracle.csr_spmv(nrow, ncol, rp, ci, va, x, y0 if accumulate else np.zeros(nrow))
        # same kernels and per-row order as the device path -> identical bits to it ...
        Y = H.Vector(y0); H.CSRMatrixMatVector(A, H.Vector(x), Y, accumulate)
        assert_bits(yh.numpy(), host(Y.values), f"host path vs device path {which} acc={accumulate}")
        # ... and the oracle's bits when the plan chose an in-order kernel, its tolerance otherwise
        if A.plan_kernel()[0] in ("stream", "scalar"):
            assert_bits(yh.numpy(), ref, f"host path {which} acc={accumulate}")
        else:
            assert max_row_error(yh.numpy(), ref, row_scale_csr(nrow, rp, ci, va, x), y0 if accumulate else None) <= TOL64
    # the launch counter counts graph replays kernel by kernel and does not count the capture itself
    assert 9 <= thsp.lib.launch_count() - launches0 < 1000


@pytest.mark.parametrize("nrow_pad", [0, 1, 2, 3])
def test_ell_fp32(thsp, cuda, oracle, nrow_pad):
    """thsp_ell_spmv_f32 (the fp32 extension of ELLMatrixMatVector, src/mat_vec.cpp:97-121): the slot-by-slot order in
    float is kept, so the result equals the float restatement bit for bit - with nrow a multiple of 4 (128-bit loads,
    four rows per thread) and not (one row per thread) - and stays within 1e-5 of the fp64 sum of the rounded inputs."""
    import ctypes as C
    from arm_spmv_b200.lib import check, current_stream, load, ptr
    nrow0, ncol, rp, ci, va = synthetic(oracle, "stencil12")
    ri = np.repeat(np.arange(nrow0, dtype=np.int32), np.diff(rp))
    nrow = nrow0 - nrow_pad          # drop the last rows: 1728 is a multiple of 4, 1727/1726/1725 are not
    keep = ri < nrow
    ri, ci, va = ri[keep], ci[keep], va[keep]
    k, eco, eva, _ = oracle.coo2ell(nrow, ncol, ri, ci, va)
    x = (oracle.gen_vector(ncol, 7) - 0.5).astype(np.float32)
    y0 = oracle.gen_vector(nrow, 8).astype(np.float32)
    v32 = eva.astype(np.float32)
    want = oracle.ell_spmv_f32(nrow, k, eco, v32, x, y0)
    y, d_col, d_val, d_x = dev(y0), dev(eco), dev(v32), dev(x)   # named: the tensors must outlive the launch
    check(load().thsp_ell_spmv_f32(nrow, ncol, k, ptr(d_col), ptr(d_val), ptr(d_x), ptr(y), current_stream()))
    assert_bits(host(y), want, "ELL fp32")
    ref64 = oracle.ell_spmv(nrow, ncol, k, eco, v32.astype(np.float64), x.astype(np.float64), y0.astype(np.float64))
    scale = row_scale_coo(nrow, ri, ci, va.astype(np.float32).astype(np.float64), x.astype(np.float64))
    assert max_row_error(host(y).astype(np.float64), ref64, scale, y0.astype(np.float64)) <= TOL32


def test_csr_host_buffer_flow_form(thsp, cuda, oracle, monkeypatch):
    """thsp_csr_plan_spmv_host_f64, flow form (stream kernel, y = A x, page-locked x and y, >= 2 M rows): one upload of x,
    one launch that multiplies behind the arriving x and stores y into the caller's vector.  Same bits as the device path
    and the oracle, call after call with changing x; an x that contains the NaN pattern the form marks missing values
    with makes it give up (here after 5 ms instead of 4 s) and the chunked form answers - same bits again."""
    import ctypes
    from arm_spmv_b200 import host as H
    n = 128
    nrow = n ** 3
    A = H.stencil27_csr(n)
    assert A.plan_kernel()[0] == "stream"
    rp, ci, va = oracle.gen_stencil27_csr(n)
    lib = thsp.load()
    xd = torch.full((nrow,), float("nan"), dtype=torch.float64, device="cuda")
    yd = torch.full((nrow,), float("nan"), dtype=torch.float64, device="cuda")
    xh = torch.empty(nrow, dtype=torch.float64).pin_memory()
    yh = torch.empty(nrow, dtype=torch.float64).pin_memory()
    for it in range(4):
        x = oracle.gen_vector(nrow, 50 + it) - 0.25 * it
        xh.copy_(torch.from_numpy(x))
        yh.fill_(float("nan"))
        l0 = thsp.lib.launch_count()
        thsp.lib.check(lib.thsp_csr_plan_spmv_host_f64(A.plan(), ctypes.c_void_p(xh.data_ptr()), ctypes.c_void_p(yh.data_ptr()),
                                                       thsp.lib.ptr(xd), thsp.lib.ptr(yd), 0, thsp.lib.current_stream()))
        assert thsp.lib.launch_count() - l0 == (3 if it == 0 else 2), "the flow form is one fill and one SpMV launch (+ the first call's look at the columns)"
        assert_bits(yh.numpy(), oracle.csr_spmv(nrow, nrow, rp, ci, va, x, np.zeros(nrow)), f"flow form, call {it}")
    # x holds the pattern itself: the flow form cannot tell it from "not arrived", gives up, the chunked form runs
    monkeypatch.setenv("THSP_FLOW_SPIN_CYCLES", "10000000")
    x = oracle.gen_vector(nrow, 77)
    x[nrow // 3] = np.frombuffer(np.uint64(0x7FF85EEDC0DEF00D).tobytes(), dtype=np.float64)[0]
    xh.copy_(torch.from_numpy(x))
    yh.fill_(0.0)
    thsp.lib.check(lib.thsp_csr_plan_spmv_host_f64(A.plan(), ctypes.c_void_p(xh.data_ptr()), ctypes.c_void_p(yh.data_ptr()),
                                                   thsp.lib.ptr(xd), thsp.lib.ptr(yd), 0, thsp.lib.current_stream()))
    Y = H.Vector(np.zeros(nrow)); H.CSRMatrixMatVector(A, H.Vector(x), Y, False)
    got, want = yh.numpy(), host(Y.values)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got).sum() == 27
    assert_bits(got[~np.isnan(got)], want[~np.isnan(want)], "chunked form after the flow form gave up")
