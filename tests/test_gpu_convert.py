"""-m gpu: format conversions on the GPU are bit-identical to the reference constructors."""
import ctypes

import numpy as np
import pytest
import torch

import cases as C
from conftest import load_golden
from gpu_util import assert_bits, dev, host

pytestmark = pytest.mark.gpu
NAMES = list(C.cases().keys())


def convert_all(H, nrow, ncol, ri, ci, va):
    A = H.COOMatrix(nrow, ncol, ri, ci, va)
    B = H.CSRMatrix(A); Cc = H.CSCMatrix(A); D = H.ELLMatrix(A)
    return A, B, Cc, D


@pytest.mark.parametrize("name", NAMES)
def test_conversions_match_golden(thsp, cuda, name):
    from arm_spmv_b200 import host as H
    g = load_golden(name)
    nrow, ncol = int(g["nrow"]), int(g["ncol"])
    A, B, Cc, D = convert_all(H, nrow, ncol, g["ri"], g["ci"], g["va"])
    assert_bits(host(B.row_ptr), g["csr_row_ptr"], "row_ptr"); assert_bits(host(B.col_ind), g["csr_col_ind"], "csr col")
    assert_bits(host(B.values), g["csr_values"], "csr val")
    assert B.ndiag == len(g["csr_diagonal"]); assert_bits(host(B.diagonal)[:B.ndiag], g["csr_diagonal"], "diag")
    assert_bits(host(Cc.col_ptr), g["csc_col_ptr"], "col_ptr"); assert_bits(host(Cc.row_ind), g["csc_row_ind"], "csc row")
    assert_bits(host(Cc.values), g["csc_values"], "csc val")
    assert D.nonzeros_in_row == int(g["ell_width"])
    assert_bits(host(D.col_ind), g["ell_col_ind"], "ell col"); assert_bits(host(D.values), g["ell_values"], "ell val")
    assert_bits(host(D.diagonal)[:D.ndiag], g["ell_diagonal"], "ell diag")
    if "dia_offsets" in g:
        E = H.DIAMatrix(B)
        assert E.ndiags == len(g["dia_offsets"])
        assert_bits(host(E.offsets), g["dia_offsets"], "dia off"); assert_bits(host(E.values), g["dia_values"], "dia val")


@pytest.mark.parametrize("kind", ["uniform", "rmat", "lap5_sorted", "stencil_shuffled"])
def test_conversions_match_oracle_medium(thsp, cuda, oracle, kind):
    """Sizes that exercise several radix passes, many CTAs and the multi-level scan."""
    from arm_spmv_b200 import host as H
    if kind == "uniform":
        nrow, ncol = 70001, 65537
        ri, ci, va = oracle.gen_uniform_coo(nrow, ncol, 1_200_003, 43)
    elif kind == "rmat":
        nrow = ncol = 1 << 17
        ri, ci, va = oracle.gen_rmat_coo(17, 1_500_000, 42)
    elif kind == "lap5_sorted":
        nrow = ncol = 300 * 300
        ri, ci, va = oracle.gen_lap5_coo(300)
    else:
        rp, cc, vv = oracle.gen_stencil27_csr(40)
        nrow = ncol = 64000
        ri = np.repeat(np.arange(nrow, dtype=np.int32), np.diff(rp))
        p = np.random.RandomState(2).permutation(len(vv))
        ri, ci, va = ri[p], cc[p], vv[p]
    if kind == "rmat":   # the ELL slab of a power-law matrix is huge: CSR/CSC only
        A = H.COOMatrix(nrow, ncol, ri, ci, va); B = H.CSRMatrix(A); Cc = H.CSCMatrix(A); D = None
    else:
        A, B, Cc, D = convert_all(H, nrow, ncol, ri, ci, va)
    rp, co, cv, dg = oracle.coo2csr(nrow, ncol, ri, ci, va)
    assert_bits(host(B.row_ptr), rp, "row_ptr"); assert_bits(host(B.col_ind), co, "csr col"); assert_bits(host(B.values), cv, "csr val")
    assert B.ndiag == len(dg); assert_bits(host(B.diagonal)[:B.ndiag], dg, "diag")
    cp, ro, cv2 = oracle.coo2csc(nrow, ncol, ri, ci, va)
    assert_bits(host(Cc.col_ptr), cp, "col_ptr"); assert_bits(host(Cc.row_ind), ro, "csc row"); assert_bits(host(Cc.values), cv2, "csc val")
    if D is not None:
        k, eco, eva, edg = oracle.coo2ell(nrow, ncol, ri, ci, va)
        assert D.nonzeros_in_row == k
        assert_bits(host(D.col_ind), eco, "ell col"); assert_bits(host(D.values), eva, "ell val")
    if kind == "lap5_sorted":
        E = H.DIAMatrix(B); off, dv = oracle.csr2dia(nrow, ncol, rp, co, cv)
        assert_bits(host(E.offsets), off, "dia off"); assert_bits(host(E.values), dv, "dia val")


@pytest.mark.parametrize("case", ["one_entry", "one_row_hub", "one_column_hub", "tile_minus_1", "tile_exact", "tile_plus_1", "two_tiles",
                                  "one_pass", "two_passes", "three_passes", "four_passes", "sorted_with_huge_gaps", "descending",
                                  "all_duplicates", "empty"])
def test_conversion_edges(thsp, cuda, oracle, case):
    """The radix sort's corner cases against the oracle's counting sort (src/matrix.cpp:125-144), bit for bit: tile
    boundaries (4096 entries per CTA), 1 to 4 digit passes (up to 2^25 buckets), hub buckets, long runs of empty
    buckets in front of / between / behind the keys, already sorted and reversed input, nothing at all."""
    from arm_spmv_b200 import host as H
    import zlib
    rs = np.random.RandomState(zlib.crc32(case.encode()))
    T = 4096
    nnz, nrow, ncol = 3 * T + 5, 5000, 4000
    if case == "one_entry":
        nnz = 1
    elif case in ("tile_minus_1", "tile_exact", "tile_plus_1", "two_tiles"):
        nnz = {"tile_minus_1": T - 1, "tile_exact": T, "tile_plus_1": T + 1, "two_tiles": 2 * T}[case]
    elif case in ("one_pass", "two_passes", "three_passes", "four_passes"):
        nrow = {"one_pass": 200, "two_passes": 60000, "three_passes": (1 << 17) + 3, "four_passes": (1 << 25) - 7}[case]
        ncol = 300
    elif case == "empty":
        nnz = 0
    ri = rs.randint(0, nrow, nnz).astype(np.int32)
    ci = rs.randint(0, ncol, nnz).astype(np.int32)
    va = rs.uniform(-1, 1, nnz)
    if case == "one_row_hub":
        ri[:] = 1234
    elif case == "one_column_hub":
        ci[:] = 77
    elif case == "sorted_with_huge_gaps":
        nrow = 1 << 22
        ri = np.sort(rs.choice(np.array([5, 6, 7, 100000, 100001, 3000000, nrow - 1], dtype=np.int32), nnz)).astype(np.int32)
    elif case == "descending":
        ri = np.sort(ri)[::-1].copy()
    elif case == "all_duplicates":
        ri[:] = 3; ci[:] = 3
    A = H.COOMatrix(nrow, ncol, ri, ci, va)
    B = H.CSRMatrix(A); Cc = H.CSCMatrix(A)
    rp, co, cv, dg = oracle.coo2csr(nrow, ncol, ri, ci, va)
    assert_bits(host(B.row_ptr), rp, "row_ptr"); assert_bits(host(B.col_ind), co, "csr col"); assert_bits(host(B.values), cv, "csr val")
    nd = min(len(dg), nrow)
    assert_bits(host(B.diagonal)[:nd], dg[:nd], "diag")
    cp, ro, cv2 = oracle.coo2csc(nrow, ncol, ri, ci, va)
    assert_bits(host(Cc.col_ptr), cp, "col_ptr"); assert_bits(host(Cc.row_ind), ro, "csc row"); assert_bits(host(Cc.values), cv2, "csc val")
    if case not in ("one_row_hub", "all_duplicates", "four_passes", "sorted_with_huge_gaps"):   # keep the ELL slab small
        D = H.ELLMatrix(A)
        k, eco, eva, _ = oracle.coo2ell(nrow, ncol, ri, ci, va)
        assert D.nonzeros_in_row == k
        assert_bits(host(D.col_ind), eco, "ell col"); assert_bits(host(D.values), eva, "ell val")


@pytest.mark.parametrize("n", [0, 1, 2, 1023, 1024, 1025, 5000, 1024 * 1024 + 17, 3_000_001])
def test_exclusive_scan(thsp, cuda, n):
    rs = np.random.RandomState(n % 97)
    a = rs.randint(0, 50, n).astype(np.int32)
    d = dev(np.concatenate([a, [0]]).astype(np.int32))
    out = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    thsp.lib.check(thsp.load().thsp_exclusive_scan_i32(n, thsp.lib.ptr(d), thsp.lib.ptr(out), thsp.lib.current_stream()))
    want = np.concatenate([[0], np.cumsum(a)]).astype(np.int32)
    assert_bits(host(out), want, f"scan {n}")
    # in place (the way row_ptr is built)
    thsp.lib.check(thsp.load().thsp_exclusive_scan_i32(n, thsp.lib.ptr(d), thsp.lib.ptr(d), thsp.lib.current_stream()))
    assert_bits(host(d), want, f"scan in place {n}")


@pytest.mark.parametrize("between", ["nothing", "other_conversion", "merge_spmv", "scratch_release", "other_width", "other_arrays",
                                     "no_prepare"])
def test_ell_two_step_conversion(thsp, cuda, oracle, between):
    """thsp_coo2ell_prepare keeps the row sort in the library's scratch for the thsp_coo2ell that follows; whatever
    happens in between (another conversion, a kernel that borrows the same scratch, different arguments, no prepare at
    all) the slab must equal the reference constructor's (src/matrix.cpp:450-500) bit for bit."""
    from arm_spmv_b200 import host as H
    lib = thsp.load(); chk = thsp.lib.check; ptr = thsp.lib.ptr; cs = thsp.lib.current_stream
    rs = np.random.RandomState(5)
    nrow, ncol, nnz = 20000, 15000, 150000
    ri = rs.randint(0, nrow, nnz).astype(np.int32); ci = rs.randint(0, ncol, nnz).astype(np.int32); va = rs.uniform(-1, 1, nnz)
    k, eco, eva, _ = oracle.coo2ell(nrow, ncol, ri, ci, va)
    A = H.COOMatrix(nrow, ncol, ri, ci, va)
    w = ctypes.c_int(0)
    if between != "no_prepare":
        chk(lib.thsp_coo2ell_prepare(nrow, ncol, nnz, ptr(A.row_ind), ptr(A.col_ind), ptr(A.values), ctypes.byref(w), cs()))
        assert w.value == k
    src = A
    if between == "other_conversion":      # overwrites the sorted entries and the row pointers in scratch
        B = H.COOMatrix(nrow, ncol, ci % nrow, ri % ncol, -va)
        H.CSRMatrix(B); H.ELLMatrix(B)
    elif between == "merge_spmv":          # the merge-path kernel's carries live in the row-pointer slot
        R = H.CSRMatrix(H.rmat_coo(12, 40000, 3))
        x = H.gen_vector(R.ncol, 1); y = H.Vector(R.nrow); y.Fill(0.0)
        H.csr_spmv_kernel(4, 1, R, x.values, y.values, True)
    elif between == "scratch_release":     # the buffers the prepare step filled are gone
        chk(lib.thsp_scratch_release())
    elif between == "other_arrays":        # same contents at other addresses: nothing to reuse
        src = H.COOMatrix(nrow, ncol, ri.copy(), ci.copy(), va.copy())
    width = k + 3 if between == "other_width" else k
    oc = torch.full((nrow * width,), -7, dtype=torch.int32, device="cuda")
    ov = torch.full((nrow * width,), float("nan"), dtype=torch.float64, device="cuda")
    chk(lib.thsp_coo2ell(nrow, ncol, nnz, ptr(src.row_ind), ptr(src.col_ind), ptr(src.values), width, ptr(oc), ptr(ov), None, None, cs()))
    if between == "other_width":   # a wider slab = the same slots followed by padding columns (col 0, +0.0)
        eco = np.concatenate([eco, np.zeros(3 * nrow, np.int32)]); eva = np.concatenate([eva, np.zeros(3 * nrow)])
    assert_bits(host(oc), eco, f"ell col ({between})"); assert_bits(host(ov), eva, f"ell val ({between})")


@pytest.mark.parametrize("between", ["nothing", "other_count", "scratch_release", "other_conversion"])
@pytest.mark.parametrize("shape", ["banded", "many_diagonals", "hub_row"])
def test_dia_count_then_emit(thsp, cuda, oracle, between, shape):
    """DIAMatrix(const CSRMatrix&) (src/matrix.cpp:673-726) as the library does it: count the diagonals, allocate, emit
    the offsets (reusing the marks of the counting call when nothing touched them in between), fill.  Banded rows go
    through the shared-memory fill, > 152 diagonals through the plain one, a hub row past the staged entries."""
    from arm_spmv_b200 import host as H
    lib = thsp.load(); chk = thsp.lib.check; ptr = thsp.lib.ptr; cs = thsp.lib.current_stream
    rs = np.random.RandomState(11)
    if shape == "banded":
        nrow = ncol = 6000
        ri = np.repeat(np.arange(nrow, dtype=np.int32), 7); ci = (ri + rs.randint(-9, 10, ri.size)).clip(0, ncol - 1).astype(np.int32)
    elif shape == "many_diagonals":
        nrow, ncol = 700, 650
        ri = rs.randint(0, nrow, 9000).astype(np.int32); ci = rs.randint(0, ncol, 9000).astype(np.int32)
    else:   # one row with 6000 entries (more than a CTA stages), duplicates included, others short
        nrow = ncol = 3000
        ri = np.concatenate([np.full(6000, 5, np.int32), rs.randint(0, nrow, 4000).astype(np.int32)])
        ci = np.concatenate([(5 + rs.randint(-60, 61, 6000)).clip(0, ncol - 1), (ri[6000:] + rs.randint(-60, 61, 4000)).clip(0, ncol - 1)]).astype(np.int32)
    keep = ~((ri == 0) & (ci == ncol - 1))
    ri, ci = ri[keep], ci[keep]
    va = rs.uniform(-1, 1, ri.size)
    rp, co, cv, _ = oracle.coo2csr(nrow, ncol, ri, ci, va)
    off, dv = oracle.csr2dia(nrow, ncol, rp, co, cv)
    d_rp, d_co, d_cv = dev(rp), dev(co), dev(cv)
    nd = ctypes.c_int(0)
    chk(lib.thsp_csr2dia_offsets(nrow, ncol, ptr(d_rp), ptr(d_co), ctypes.byref(nd), None, 0, cs()))
    assert nd.value == len(off)
    if between == "other_count":
        o_rp, o_co = dev(np.array([0, 1, 2, 2], np.int32)), dev(np.array([0, 2], np.int32))   # (0,0) and (1,2): offsets 0, +1
        k = ctypes.c_int(0)
        chk(lib.thsp_csr2dia_offsets(3, 3, ptr(o_rp), ptr(o_co), ctypes.byref(k), None, 0, cs()))
        assert k.value == 2
    elif between == "scratch_release":
        chk(lib.thsp_scratch_release())
    elif between == "other_conversion":
        H.CSCMatrix(H.COOMatrix(nrow, ncol, ri, ci, va)); H.ELLMatrix(H.COOMatrix(50, 50, ri[:40] % 50, ci[:40] % 50, va[:40]))
    d_off = torch.full((nd.value,), -12345, dtype=torch.int32, device="cuda")
    d_val = torch.full((nrow * nd.value,), float("nan"), dtype=torch.float64, device="cuda")
    chk(lib.thsp_csr2dia_offsets(nrow, ncol, ptr(d_rp), ptr(d_co), ctypes.byref(nd), ptr(d_off), nd.value, cs()))
    assert nd.value == len(off)
    chk(lib.thsp_csr2dia_fill(nrow, ncol, ptr(d_rp), ptr(d_co), ptr(d_cv), nd.value, ptr(d_off), ptr(d_val), cs()))
    assert_bits(host(d_off), off, f"dia offsets ({shape}, {between})"); assert_bits(host(d_val), dv, f"dia values ({shape}, {between})")


@pytest.mark.parametrize("hub", [0, 4999, 9000])
def test_ell_long_rows(thsp, cuda, oracle, hub):
    """ELLMatrix(const COOMatrix&) when rows average >= 20 entries (the kernel that stages a CTA's entries in shared
    memory), without and with one row longer than the stage (4096 entries)."""
    from arm_spmv_b200 import host as H
    rs = np.random.RandomState(hub + 1)
    nrow, ncol = 300, 12000
    ri = np.repeat(np.arange(nrow, dtype=np.int32), 24)
    if hub:
        ri = np.concatenate([ri, np.full(hub, 137, np.int32)])
    rs.shuffle(ri)
    ci = rs.randint(0, ncol, ri.size).astype(np.int32); va = rs.uniform(-1, 1, ri.size)
    D = H.ELLMatrix(H.COOMatrix(nrow, ncol, ri, ci, va))
    k, eco, eva, _ = oracle.coo2ell(nrow, ncol, ri, ci, va)
    assert D.nonzeros_in_row == k == 24 + hub
    assert_bits(host(D.col_ind), eco, "ell col"); assert_bits(host(D.values), eva, "ell val")


def test_generators_match_cpu_twins(thsp, cuda, oracle):
    from arm_spmv_b200 import host as H
    A = H.stencil27_csr(9)
    rp, ci, va = oracle.gen_stencil27_csr(9)
    assert_bits(host(A.row_ptr), rp, "st rp"); assert_bits(host(A.col_ind), ci, "st ci"); assert_bits(host(A.values), va, "st va")
    S = H.stencil27_csr(9, 100, 517)
    rp2, ci2, va2 = oracle.gen_stencil27_csr(9, 100, 517)
    assert_bits(host(S.row_ptr), rp2, "slab rp"); assert_bits(host(S.col_ind), ci2, "slab ci")
    E = H.stencil27_ell(9)
    ri = np.repeat(np.arange(729, dtype=np.int32), np.diff(rp))
    k, eco, eva, _ = oracle.coo2ell(729, 729, ri, ci, va)
    assert k == 27
    assert_bits(host(E.col_ind), eco, "ell gen col"); assert_bits(host(E.values), eva, "ell gen val")
    Cq = H.stencil27_coo(9)
    assert_bits(host(Cq.row_ind), ri, "coo gen row"); assert_bits(host(Cq.col_ind), ci, "coo gen col")
    Lp = H.lap5_coo(33); lri, lci, lva = oracle.gen_lap5_coo(33)
    assert_bits(host(Lp.row_ind), lri, "lap ri"); assert_bits(host(Lp.col_ind), lci, "lap ci"); assert_bits(host(Lp.values), lva, "lap va")
    U = H.uniform_coo(5000, 4000, 100003, 43); uri, uci, uva = oracle.gen_uniform_coo(5000, 4000, 100003, 43)
    assert_bits(host(U.row_ind), uri, "uni ri"); assert_bits(host(U.col_ind), uci, "uni ci"); assert_bits(host(U.values), uva, "uni va")
    Rm = H.rmat_coo(14, 50000, 42); rri, rci, rva = oracle.gen_rmat_coo(14, 50000, 42)
    assert_bits(host(Rm.row_ind), rri, "rmat ri"); assert_bits(host(Rm.col_ind), rci, "rmat ci"); assert_bits(host(Rm.values), rva, "rmat va")
    assert_bits(host(H.gen_vector(10007, 5).values), oracle.gen_vector(10007, 5), "vec")


def _sorted_by_row_no_dups(rs, nrow, ncol, nnz):
    """nnz distinct (row, col) pairs in row order, the columns of a row in a shuffled order."""
    lin = np.unique(rs.randint(0, nrow * ncol, nnz, dtype=np.int64))
    rs.shuffle(lin)
    ri = (lin // ncol).astype(np.int32); ci = (lin % ncol).astype(np.int32)
    o = np.argsort(ri, kind="stable")
    return ri[o], ci[o], rs.uniform(-1, 1, len(o))


@pytest.mark.parametrize("case", ["lap5", "stencil27", "shuffled_columns", "column_by_column_to_csr", "tails", "one_row",
                                  "stage_overflow", "length_64", "length_65", "duplicates", "order_breaks_late",
                                  "index_out_of_range_free", "empty_buckets", "too_wide_for_l2"])
def test_transposed_conversions(thsp, cuda, oracle, case):
    """COO -> CSC of entries that come row by row (and COO -> CSR of entries that come column by column) skips the radix
    sort: per-bucket cursors + a per-bucket sort by entry number (convert.cu, transpose_entries).  Same arrays as the
    reference's counting sort (src/matrix.cpp:295-325 / 115-154) bit for bit, whichever way the call went; the way it
    went is checked too (thsp_coo_last_path: 2 = transposed, 3 = tried and handed to the sort)."""
    from arm_spmv_b200 import host as H
    import zlib
    lib = thsp.load()
    rs = np.random.RandomState(zlib.crc32(case.encode()))
    to_csr = False
    want_path = 2
    if case == "lap5":
        nrow = ncol = 257 * 257
        ri, ci, va = oracle.gen_lap5_coo(257)
    elif case == "stencil27":
        rp, cc, vv = oracle.gen_stencil27_csr(33)
        nrow = ncol = 33 ** 3
        ri = np.repeat(np.arange(nrow, dtype=np.int32), np.diff(rp)); ci = cc; va = rs.uniform(-1, 1, len(vv))   # not symmetric
    elif case == "shuffled_columns":
        nrow, ncol = 50021, 70001
        ri, ci, va = _sorted_by_row_no_dups(rs, nrow, ncol, 900_000)
    elif case == "column_by_column_to_csr":
        ncol, nrow = 50021, 70001
        ci, ri, va = _sorted_by_row_no_dups(rs, ncol, nrow, 900_000)
        to_csr = True
    elif case == "tails":   # entry counts around the four-entry groups of the count / place kernels
        for nnz in (1, 2, 3, 5, 6, 7, 1023, 1025, 4099):
            nrow, ncol = 301, 211
            ri, ci, va = _sorted_by_row_no_dups(rs, nrow, ncol, nnz)
            Cc = H.CSCMatrix(H.COOMatrix(nrow, ncol, ri, ci, va))
            cp, ro, cv = oracle.coo2csc(nrow, ncol, ri, ci, va)
            assert_bits(host(Cc.col_ptr), cp, f"col_ptr {nnz}"); assert_bits(host(Cc.row_ind), ro, f"csc row {nnz}")
            assert_bits(host(Cc.values), cv, f"csc val {nnz}")
        return
    elif case == "one_row":
        nrow, ncol = 1, 5000
        ci = rs.permutation(ncol).astype(np.int32)[:3001]; ri = np.zeros(len(ci), np.int32); va = rs.uniform(-1, 1, len(ci))
    elif case in ("stage_overflow", "length_64", "length_65"):
        # 300 neighbouring columns of L entries each in a wide, otherwise nearly empty matrix: the mean bucket is short, so a
        # CTA takes 128 buckets, and 128 * L entries do not fit its 4096-entry stage: the later ones are sorted in place
        L = {"stage_overflow": 60, "length_64": 64, "length_65": 65}[case]
        nrow, ncol = L + 7, 200_000
        cols = 1000 + np.arange(300, dtype=np.int32)
        ri = np.repeat(np.arange(L, dtype=np.int32), len(cols))
        ci = np.concatenate([rs.permutation(cols) for _ in range(L)]).astype(np.int32)
        extra_r = np.full(50, L + 3, np.int32); extra_c = rs.choice(np.arange(5000, 190_000, dtype=np.int32), 50, replace=False)
        ri = np.concatenate([ri, extra_r]); ci = np.concatenate([ci, extra_c]); va = rs.uniform(-1, 1, len(ri))
        if case == "length_65":
            want_path = 3
    elif case == "duplicates":
        nrow, ncol = 5003, 4001
        ri, ci, va = _sorted_by_row_no_dups(rs, nrow, ncol, 60_000)
        k = len(ri) // 2   # one pair of entries with both indices equal (they sit in the same row, so next to each other)
        ri = np.insert(ri, k + 1, ri[k]); ci = np.insert(ci, k + 1, ci[k]); va = np.insert(va, k + 1, 0.5)
        for q in range(0, len(ri) - 3, 1000):   # and whole runs of them here and there: they must keep their COO order
            ri[q:q + 3] = ri[q]; ci[q:q + 3] = ci[q]
    elif case == "order_breaks_late":
        nrow, ncol = 200_003, 150_001
        ri, ci, va = _sorted_by_row_no_dups(rs, nrow, ncol, 1_300_000)
        assert len(ri) > (1 << 20) + 1000
        ri[-7], ri[-400] = ri[-400], ri[-7]   # past the first 2^20 entries, which is all the probe reads
        want_path = 3
    elif case == "index_out_of_range_free":   # the largest legal indices
        nrow, ncol = 77, 91
        ri, ci, va = _sorted_by_row_no_dups(rs, nrow, ncol, 3000)
        ri[-1] = nrow - 1; ci[-1] = ncol - 1; ci[0] = 0
        lin = ri.astype(np.int64) * ncol + ci
        if len(np.unique(lin)) != len(lin):
            want_path = None
    elif case == "too_wide_for_l2":   # row by row, but a column's entries come from all over the matrix: the probe sends it to the sort
        nrow = ncol = 1 << 22
        ri, ci, va = _sorted_by_row_no_dups(rs, nrow, ncol, 3_000_000)
        want_path = 1
    else:   # empty_buckets: long stretches of empty columns in front of, between and behind the occupied ones
        nrow, ncol = 4000, 1 << 22
        ri, ci, va = _sorted_by_row_no_dups(rs, nrow, 1024, 20_000)
        ci = (ci.astype(np.int64) * 4001 + 100_000).astype(np.int32)
        assert int(ci.max()) < ncol
    A = H.COOMatrix(nrow, ncol, ri, ci, va)
    if to_csr:
        B = H.CSRMatrix(A)
        path = lib.thsp_coo_last_path()
        rp, co, cv, dg = oracle.coo2csr(nrow, ncol, ri, ci, va)
        assert_bits(host(B.row_ptr), rp, "row_ptr"); assert_bits(host(B.col_ind), co, "csr col"); assert_bits(host(B.values), cv, "csr val")
        assert B.ndiag == len(dg); assert_bits(host(B.diagonal)[:B.ndiag], dg, "diag")
    else:
        Cc = H.CSCMatrix(A)
        path = lib.thsp_coo_last_path()
        cp, ro, cv = oracle.coo2csc(nrow, ncol, ri, ci, va)
        assert_bits(host(Cc.col_ptr), cp, "col_ptr"); assert_bits(host(Cc.row_ind), ro, "csc row"); assert_bits(host(Cc.values), cv, "csc val")
    if want_path is not None:
        assert path == want_path, f"conversion went way {path}, expected {want_path}"
    # the same matrix the other way round keeps its keys: copied through (path 0)
    if not to_csr:
        B = H.CSRMatrix(A)
        if case != "order_breaks_late":
            assert lib.thsp_coo_last_path() == 0
        rp, co, cv, dg = oracle.coo2csr(nrow, ncol, ri, ci, va)
        assert_bits(host(B.row_ptr), rp, "row_ptr"); assert_bits(host(B.col_ind), co, "csr col"); assert_bits(host(B.values), cv, "csr val")
