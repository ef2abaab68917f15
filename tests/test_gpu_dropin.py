"""-m gpu: the C++ drop-in layer.  bin/api_check (tests/cpp/api_check.cpp linked against include/ +
bin/TH_sparse.a) is run on the GPU and its dumps are compared with the dumps the SAME source
produced when built against the unmodified reference (fixtures in tests/golden/api_*/, made by
tests/golden/make_api_golden.py).  bin/main is the reference's own main.cpp linked against this
library with the reference's link line."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXACT = ["csr_row_ptr.i32", "csr_col_ind.i32", "csr_values.f64", "csc_col_ptr.i32", "csc_row_ind.i32", "csc_values.f64",
         "ell_col_ind.i32", "ell_values.f64", "dia_offsets.i32", "dia_values.f64", "meta.i32", "x.f64",
         "y_csr.f64", "y_ell.f64", "y_dia.f64", "y_csr_copy.f64", "y_csr_assign.f64", "w_axpby.f64"]
CLOSE = ["y_coo.f64", "y_csc.f64", "w_normalised.f64", "z_chain.f64"]


def load(path):
    return np.fromfile(path, dtype=np.int32 if path.endswith(".i32") else np.float64)


@pytest.mark.parametrize("name", ["lap5_12", "rand90"])
def test_cpp_api_matches_reference_build(name, tmp_path):
    exe = os.path.join(ROOT, "bin", "api_check")
    assert os.path.exists(exe), "bin/api_check missing: run __graft_entry__.build()"
    gold = os.path.join(ROOT, "tests", "golden", "api_" + name)
    r = subprocess.run([exe, os.path.join(gold, "matrix.mtx"), str(tmp_path), "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for f in EXACT:
        a, b = load(os.path.join(str(tmp_path), f)), load(os.path.join(gold, f))
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), f"{name}/{f} differs from the reference build"
    for f in CLOSE:
        a, b = load(os.path.join(str(tmp_path), f)), load(os.path.join(gold, f))
        assert a.shape == b.shape
        assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, float(np.max(np.abs(b)))), f
    s, g = load(os.path.join(str(tmp_path), "scalars.f64")), load(os.path.join(gold, "scalars.f64"))
    assert abs(s[0] - g[0]) <= 1e-12 * abs(g[0]) and abs(s[1] - g[1]) <= 1e-12
    assert s[2] == g[2] == 1.0 and s[3] == g[3] == 0.0
    # partitioned variants: 50 accumulations of A x, written back into y (this build only)
    one = load(os.path.join(str(tmp_path), "y_csr_copy.f64"))
    acc = np.zeros_like(one)
    for _ in range(50):
        acc = acc + one
    got = load(os.path.join(str(tmp_path), "y_csr_numa.f64"))
    assert got.tobytes() == acc.tobytes(), "CSR partitioned result != 50 in-order accumulations"
    for f in ("y_ell_numa.f64", "y_coo_numa.f64", "y_csc_numa.f64"):
        got = load(os.path.join(str(tmp_path), f))
        assert np.max(np.abs(got - acc)) <= 1e-11 * max(1.0, float(np.max(np.abs(acc)))), f
    dia1 = load(os.path.join(str(tmp_path), "y_dia.f64")) / 3.0
    got = load(os.path.join(str(tmp_path), "y_dia_numa.f64"))
    assert np.max(np.abs(got - 50 * dia1)) <= 1e-11 * max(1.0, float(np.max(np.abs(50 * dia1))))
    for line in ("### CSR NUMA GFLOPS", "### ELL NUMA GFLOPS", "### COO NUMA GFLOPS", "### CSC NUMA GFLOPS", "### DIA NUMA GFLOPS"):
        assert line in r.stdout


def test_reference_main_runs_on_the_gpu_library(tmp_path):
    """The reference's unmodified driver: argv = <file.mtx> <nthreads>, ten '###' GFLOPS lines."""
    exe = os.path.join(ROOT, "bin", "main")
    if not os.path.exists(exe):
        pytest.skip("bin/main is linked from /root/reference/main.cpp in the build container only")
    mtx = os.path.join(ROOT, "tests", "golden", "api_lap5_12", "matrix.mtx")
    r = subprocess.run([exe, mtx, "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "### ROW=144, COL=144, NNZ=672" in r.stdout
    for fmt in ("COO", "CSR", "CSC", "ELL", "DIA"):
        for kind in ("CPU", "NUMA"):   # the reference's label for its un-partitioned loop is "CPU"
            m = re.search(rf"### {fmt} {kind} GFLOPS = ([0-9.eE+-]+|inf|nan)", r.stdout)
            assert m, f"missing '### {fmt} {kind} GFLOPS' in:\n{r.stdout}"


def test_adopted_arrays_are_mirrored_once_and_never_stale(tmp_path):
    """bin/adopt_check: CSRMatrix / COOMatrix / Vector built on caller new[] arrays (the adopting constructors,
    src/matrix.cpp:12-15,88-91).  Every product equals the host loop; THSP_TRACE shows the matrix arrays uploaded by the
    first product only, again after the caller refilled them, and the plan rebuilt after row_ptr[nrow] changed."""
    exe = os.path.join(ROOT, "bin", "adopt_check")
    assert os.path.exists(exe), "bin/adopt_check missing: run __graft_entry__.build()"
    r = subprocess.run([exe, "300000"], capture_output=True, text=True, timeout=300, env=dict(os.environ, THSP_TRACE="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALL OK" in r.stdout and "MISMATCH" not in r.stdout
    calls = r.stderr.split("[adopt] call ")
    uploads = [c.count("device mirror of") for c in calls[1:]]
    assert uploads[0] == 3, uploads           # row_ptr, col_ind, values
    assert uploads[1] == 0 and uploads[2] == 0, uploads   # found again; only x and y move
    assert uploads[3] >= 1, uploads           # values refilled: uploaded again
    assert uploads[4] >= 1, uploads           # row_ptr rewritten: uploaded again, plan rebuilt for the new entry count
