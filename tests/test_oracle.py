"""Pins the CPU oracle (oracle/oracle.c) against the reference's own outputs:
the committed golden fixtures (generated from the unmodified reference by
tests/golden/make_golden.py) and, when oracle/_ref/libref.so is present, the live reference.
Integer/index outputs and every in-order floating-point result must be bit-identical."""
import numpy as np
import pytest

import cases as C
from conftest import load_golden

NAMES = list(C.cases().keys())


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert a.tobytes() == b.tobytes()


def test_known_answer_vector(oracle):
    """SURVEY.md appendix A.1, hand-checked."""
    I = [3, 1, 0, 1, 3, 1, 0, 3]; J = [4, 2, 0, 0, 3, 2, 3, 0]; V = [1, 2, 3, 4, 5, 6, 7, 8.0]
    rp, co, va, dg = oracle.coo2csr(4, 5, I, J, V)
    assert rp.tolist() == [0, 2, 5, 5, 8] and co.tolist() == [0, 3, 2, 0, 2, 4, 3, 0]
    assert va.tolist() == [3, 7, 2, 4, 6, 1, 5, 8] and dg.tolist() == [3, 5]
    cp, ro, vc = oracle.coo2csc(4, 5, I, J, V)
    assert cp.tolist() == [0, 3, 3, 5, 7, 8] and ro.tolist() == [0, 1, 3, 1, 1, 3, 0, 3] and vc.tolist() == [3, 4, 8, 2, 6, 5, 7, 1]
    k, ec, ev, _ = oracle.coo2ell(4, 5, I, J, V)
    assert k == 3 and ec.tolist() == [0, 2, 0, 4, 3, 0, 0, 3, 0, 2, 0, 0] and ev.tolist() == [3, 2, 0, 1, 7, 4, 0, 5, 0, 6, 0, 8]
    off, dv = oracle.csr2dia(4, 5, rp, co, va)
    assert off.tolist() == [-3, -1, 0, 1, 3]
    assert dv.tolist() == [0, 0, 3, 0, 7, 0, 4, 0, 6, 0, 0, 0, 0, 0, 0, 8, 0, 5, 1, 0]
    x = np.arange(1, 6.0); z = np.zeros(4)
    assert oracle.coo_spmv(4, 5, I, J, V, x, z).tolist() == [31, 28, 0, 33]
    assert oracle.csr_spmv(4, 5, rp, co, va, x, z + 10).tolist() == [41, 38, 10, 43]   # proves y +=
    assert oracle.csc_spmv(4, 5, cp, ro, vc, x, z).tolist() == [31, 28, 0, 33]
    assert oracle.ell_spmv(4, 5, k, ec, ev, x, z).tolist() == [31, 28, 0, 33]
    assert oracle.dia_spmv(4, 5, off, dv, x, z).tolist() == [31, 22, 0, 28]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(oracle, name):
    g = load_golden(name)
    nrow, ncol = int(g["nrow"]), int(g["ncol"])
    ri, ci, va, x, y0 = g["ri"], g["ci"], g["va"], g["x"], g["y0"]
    rp, co, cv, dg = oracle.coo2csr(nrow, ncol, ri, ci, va)
    eq(rp, g["csr_row_ptr"]); eq(co, g["csr_col_ind"]); eq(cv, g["csr_values"]); eq(dg, g["csr_diagonal"])
    cp, ro, cv2 = oracle.coo2csc(nrow, ncol, ri, ci, va)
    eq(cp, g["csc_col_ptr"]); eq(ro, g["csc_row_ind"]); eq(cv2, g["csc_values"])
    k, eco, eva, edg = oracle.coo2ell(nrow, ncol, ri, ci, va)
    assert k == int(g["ell_width"])
    eq(eco, g["ell_col_ind"]); eq(eva, g["ell_values"]); eq(edg, g["ell_diagonal"])
    eq(oracle.coo_spmv(nrow, ncol, ri, ci, va, x, y0), g["y_coo"])
    eq(oracle.csr_spmv(nrow, ncol, rp, co, cv, x, y0), g["y_csr"])
    eq(oracle.csc_spmv(nrow, ncol, cp, ro, cv2, x, y0), g["y_csc"])
    eq(oracle.ell_spmv(nrow, ncol, k, eco, eva, x, y0), g["y_ell"])
    if "dia_offsets" in g:
        off, dv = oracle.csr2dia(nrow, ncol, rp, co, cv)
        eq(off, g["dia_offsets"]); eq(dv, g["dia_values"])
        if "y_dia" in g:
            eq(oracle.dia_spmv(nrow, ncol, off, dv, x, y0), g["y_dia"])


def test_oracle_vector_ops_match_golden(oracle):
    g = load_golden("vec_ops")
    x, y, v = g["x"], g["y"], g["v"]
    assert oracle.dot(x, y) == float(g["dot"][0])   # one thread: same serial order
    for i, (a, b) in enumerate(C.AXPBY_COEFFS):
        eq(oracle.axpby(a, x, b, y), g[f"axpby_{i}"])
    eq(oracle.fill(17, 3.25), g["fill"])
    eq(oracle.scale(1.7, v), g["scale"])
    eq(oracle.shift(-0.3, v), g["shift"])
    for i, a in enumerate(C.ADD_SCALED_COEFFS):
        eq(oracle.add_scaled(a, x, v), g[f"add_scaled_{i}"])
    for i, (a, b) in enumerate(C.ADD2_COEFFS):
        eq(oracle.add2_scaled(a, x, b, y, v), g[f"add2_scaled_{i}"])
    assert oracle.check_vector(x, x + 5e-7) and not oracle.check_vector(x, x + 2e-6)
    assert not oracle.check_vector(x, x[:-1])


@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_live_reference(oracle, ref, seed):
    """Random shapes straight against the compiled reference (build container only)."""
    rs = np.random.RandomState(1000 + seed)
    nrow, ncol = int(rs.randint(1, 200)), int(rs.randint(1, 200))
    nnz = int(rs.randint(0, 3000))
    ri = rs.randint(0, nrow, nnz).astype(np.int32); ci = rs.randint(0, ncol, nnz).astype(np.int32)
    keep = ~((ri == 0) & (ci == ncol - 1))
    ri, ci = ri[keep], ci[keep]
    # the reference's diagonal[] holds nrow entries: drop surplus row==col duplicates
    d = np.flatnonzero(ri == ci)
    if len(d) > nrow:
        drop = np.zeros(len(ri), bool); drop[d[nrow:]] = True
        ri, ci = ri[~drop], ci[~drop]
    va = rs.uniform(-1, 1, len(ri)); x = rs.uniform(0, 1, ncol); y0 = rs.uniform(-1, 1, nrow)
    a = oracle.coo2csr(nrow, ncol, ri, ci, va); b = ref.coo2csr(nrow, ncol, ri, ci, va)
    for u, w in zip(a, b): eq(u, w)
    rp, co, cv, _ = a
    for u, w in zip(oracle.coo2csc(nrow, ncol, ri, ci, va), ref.coo2csc(nrow, ncol, ri, ci, va)): eq(u, w)
    ea = oracle.coo2ell(nrow, ncol, ri, ci, va); eb = ref.coo2ell(nrow, ncol, ri, ci, va)
    assert ea[0] == eb[0]
    for u, w in zip(ea[1:], eb[1:]): eq(u, w)
    eq(oracle.csr_spmv(nrow, ncol, rp, co, cv, x, y0), ref.csr_spmv(nrow, ncol, rp, co, cv, x, y0))
    eq(oracle.coo_spmv(nrow, ncol, ri, ci, va, x, y0), ref.coo_spmv(nrow, ncol, ri, ci, va, x, y0))
    eq(oracle.ell_spmv(nrow, ncol, ea[0], ea[1], ea[2], x, y0), ref.ell_spmv(nrow, ncol, ea[0], ea[1], ea[2], x, y0))
    da = oracle.csr2dia(nrow, ncol, rp, co, cv); db = ref.csr2dia(nrow, ncol, rp, co, cv)
    for u, w in zip(da, db): eq(u, w)


def test_rebuilt_o3_reference_is_a_timing_baseline_only(oracle):
    """oracle/_ref/libref_o3.so (the reference at -O3 -march=x86-64-v3, timed beside the stock build by bench.py) loads
    next to libref.so and computes the same SpMV up to FMA contraction: index work identical, values within 1e-12."""
    import pyoracle
    if not pyoracle.RefO3.runnable():
        pytest.skip("libref_o3.so not built or this CPU lacks AVX2/FMA")
    r3 = pyoracle.RefO3(); r3.set_threads(2)
    rs = np.random.RandomState(7)
    nrow, ncol, nnz = 300, 280, 4000
    ri = rs.randint(0, nrow, nnz).astype(np.int32); ci = rs.randint(0, ncol, nnz).astype(np.int32)
    keep = ~((ri == ci) | ((ri == 0) & (ci == ncol - 1)))
    ri, ci = ri[keep], ci[keep]
    va = rs.uniform(-1, 1, len(ri)); x = rs.uniform(0, 1, ncol); y0 = rs.uniform(-1, 1, nrow)
    a = oracle.coo2csr(nrow, ncol, ri, ci, va); b = r3.coo2csr(nrow, ncol, ri, ci, va)
    for u, w in zip(a[:3], b[:3]): eq(u, w)
    rp, co, cv, _ = a
    y, y3 = oracle.csr_spmv(nrow, ncol, rp, co, cv, x, y0), r3.csr_spmv(nrow, ncol, rp, co, cv, x, y0)
    scale = np.abs(y0) + np.bincount(ri, np.abs(va * x[ci]), nrow)
    assert np.max(np.abs(y - y3) / np.maximum(scale, 1e-300)) <= 1e-12
    assert r3.time_csr_spmv(nrow, ncol, rp, co, cv, x, 2) > 0


def test_generators_are_consistent(oracle):
    """The synthetic matrices the benches use: structure checks on the CPU twins."""
    rp, ci, va = oracle.gen_stencil27_csr(5)
    assert rp[-1] == (3 * 5 - 2) ** 3 == len(ci)
    assert np.all(np.diff(rp) >= 8) and np.all(np.diff(rp) <= 27)
    for r in (0, 17, 62, 124):
        cols = ci[rp[r]:rp[r + 1]]
        assert np.all(np.diff(cols) > 0) and r in cols
        assert va[rp[r]:rp[r + 1]][list(cols).index(r)] == 26.0
    sub_rp, sub_ci, sub_va = oracle.gen_stencil27_csr(5, 30, 77)
    assert np.array_equal(sub_rp, rp[30:78] - rp[30]) and np.array_equal(sub_ci, ci[rp[30]:rp[77]])
    ri, cj, v = oracle.gen_lap5_coo(6)
    assert len(v) == 5 * 36 - 4 * 6 and np.all(np.diff(ri) >= 0)
    ri, cj, v = oracle.gen_uniform_coo(1000, 900, 5000, 43)
    assert ri.min() >= 0 and ri.max() < 1000 and cj.max() < 900 and 0 <= v.min() and v.max() < 1
    ri, cj, v = oracle.gen_rmat_coo(10, 8000, 42)
    assert ri.max() < 1024 and cj.max() < 1024
    assert np.bincount(ri, minlength=1024).max() > 20 * 8000 / 1024   # power-law head


def test_partition_and_slice(oracle):
    """src/mat_vec.cpp:233-263: equal row blocks, remainder to the last, row_ptr rebased to 0."""
    rp, ci, va = oracle.gen_stencil27_csr(4)
    n = 64
    for parts in (1, 2, 3, 8):
        covered = 0
        for p in range(parts):
            s, c = oracle.partition(n, parts, p)
            assert s == covered
            covered += c
            sub, nnz = oracle.csr_slice(rp, s, c)
            assert sub[0] == 0 and sub[-1] == nnz == rp[s + c] - rp[s]
        assert covered == n


def test_symgs_and_cg_restatements(oracle):
    """The solver twins of oracle.c (no reference counterpart: the reference keeps `diagonal` "for SymGS" and never sweeps):
    the multicolour sweep with one row per colour IS the sequential sweep; preconditioned CG solves the stencil system and
    the SymGS preconditioner saves iterations; the canonical dot is a dot."""
    n = 7
    N = n ** 3
    rp, ci, va = oracle.gen_stencil27_csr(n)
    diag = np.full(N, 26.0)
    b = oracle.gen_vector(N, 3)
    cp, perm = np.arange(N + 1, dtype=np.int32), np.arange(N, dtype=np.int32)
    x0 = oracle.gen_vector(N, 4)
    assert oracle.symgs(cp, perm, rp, ci, va, diag, b, x0).tobytes() == oracle.symgs_sequential(rp, ci, va, diag, b, x0).tobytes()
    M = np.zeros((N, N))
    for r in range(N):
        M[r, ci[rp[r]:rp[r + 1]]] = va[rp[r]:rp[r + 1]]
    want = np.linalg.solve(M, b)
    its = {}
    for kind in (0, 1, 2):
        x, it, rel = oracle.cg(rp, ci, va, diag, b, np.zeros(N), 300, 1e-12, kind, cp, perm)
        assert rel <= 1e-12 and np.max(np.abs(x - want)) <= 1e-10 * np.max(np.abs(want))
        its[kind] = it
    assert its[2] < its[0]
    a, c = oracle.gen_vector(1000, 1) - 0.5, oracle.gen_vector(1000, 2)
    assert abs(oracle.dot_canonical(a, c) - float(np.dot(a, c))) <= 1e-13 * float(np.sum(np.abs(a * c)))
    y = oracle.gen_vector(777, 9)
    assert oracle.tree_sum(oracle.tile_sumsq(y)) == oracle.dot_canonical(y, y)
    assert (oracle.hash_f64(y[:300], 10) + oracle.hash_f64(y[300:], 310)) % (1 << 64) == oracle.hash_f64(y, 10)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_buckets_ordered_by_entry_number_are_the_counting_sort(oracle, seed):
    """What the GPU's transposing conversion relies on (convert.cu, transpose_entries): drop every entry into its
    column in ANY order, then order each column by entry number - that is the reference's counting sort
    (src/matrix.cpp:295-325), duplicates included.  The arrival order of the atomics is played by a random permutation."""
    rs = np.random.RandomState(seed)
    nrow, ncol, nnz = 300, 200, 5000
    ri = np.sort(rs.randint(0, nrow, nnz)).astype(np.int32)
    ci = rs.randint(0, ncol, nnz).astype(np.int32)          # (row, col) pairs repeat: 5000 draws from 60000 cells
    va = rs.uniform(-1, 1, nnz)
    cp, ro, vo = oracle.coo2csc(nrow, ncol, ri, ci, va)
    counts = np.bincount(ci, minlength=ncol)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    cursor = ptr[:-1].copy()
    slot_entry = np.full(nnz, -1, np.int64)
    for e in rs.permutation(nnz):                            # the order the atomics happen to run in
        slot_entry[cursor[ci[e]]] = e
        cursor[ci[e]] += 1
    for c in range(ncol):
        slot_entry[ptr[c]:ptr[c + 1]].sort()                 # per-bucket sort by entry number
    eq(ptr, cp); eq(ri[slot_entry], ro); eq(va[slot_entry], vo)
