"""-m gpu: the Matrix Market entry parser on the GPU (thsp_mtx_parse_coo) against Python's int()/float() -
float(str) is the correctly rounded double, the same bits as the strtod behind the reference's %lg
(src/data_io.cpp:85) - and against the golden files the unmodified reference read."""
import os
import struct

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_mtx(path, nrow, ncol, entries, sep=" ", eol="\n", header_comments=("% a comment",)):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        for c in header_comments:
            f.write(c + "\n")
        f.write(f"{nrow} {ncol} {len(entries)}\n")
        for i, j, v in entries:
            f.write(f"{i}{sep}{j}{sep}{v}{eol}")


def bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


def test_values_are_strtod_bits(thsp, cuda, tmp_path):
    from arm_spmv_b200 import host as H
    rs = np.random.RandomState(4)
    n = 20000
    vals = []
    for k in range(n):
        kind = k % 6
        if kind == 0:
            vals.append(repr(float(rs.uniform())))                      # shortest round-trip form
        elif kind == 1:
            vals.append("%.17g" % rs.uniform(-1e3, 1e3))
        elif kind == 2:
            vals.append("%.17e" % struct.unpack("d", struct.pack("Q", int(rs.randint(0, 2 ** 62))))[0])   # any exponent, subnormals
        elif kind == 3:
            vals.append(str(int(rs.randint(-50, 50))))                  # integers, as in stencil matrices
        elif kind == 4:
            vals.append("%d.%de%d" % (rs.randint(0, 10 ** 9), rs.randint(0, 10 ** 9), rs.randint(-320, 300)))
        else:
            vals.append(str(2 ** 53 + 2 * int(rs.randint(0, 2 ** 40)) + 1))   # exactly half way between two doubles
    ent = [(int(rs.randint(1, 1001)), int(rs.randint(1, 2001)), v) for v in vals]
    path = str(tmp_path / "m.mtx")
    write_mtx(path, 1000, 2000, ent)
    A = H.COOMatrixRead(path)
    assert (A.nrow, A.ncol, A.nnz) == (1000, 2000, n)
    assert np.array_equal(A.row_ind.cpu().numpy(), np.array([e[0] - 1 for e in ent], np.int32))
    assert np.array_equal(A.col_ind.cpu().numpy(), np.array([e[1] - 1 for e in ent], np.int32))
    want = np.array([float(v) for v in vals])
    got = A.values.cpu().numpy()
    bad = np.nonzero(bits(got) != bits(want))[0]
    assert bad.size == 0, (vals[bad[0]], got[bad[0]], want[bad[0]])


@pytest.mark.parametrize("sep,eol", [(" ", "\n"), ("\t", "\r\n"), ("   ", " \n"), ("\n", "\n"), (" ", " ")])
def test_whitespace_layouts(thsp, cuda, tmp_path, sep, eol):
    """fscanf reads tokens, not lines: any whitespace between the three fields of an entry, entries split over lines,
    several entries per line, no newline at the end of the file."""
    from arm_spmv_b200 import host as H
    ent = [(1 + k % 7, 1 + k % 5, "%.3f" % (k * 0.25 - 3)) for k in range(5000)]
    path = str(tmp_path / "m.mtx")
    write_mtx(path, 7, 5, ent, sep=sep, eol=eol, header_comments=("%", "% two comments", ""))
    A = H.COOMatrixRead(path)
    assert np.array_equal(A.row_ind.cpu().numpy(), np.array([e[0] - 1 for e in ent], np.int32))
    assert np.array_equal(bits(A.values.cpu().numpy()), bits([float(e[2]) for e in ent]))


@pytest.mark.parametrize("bad", ["nan", "inf", "0x1p3", "1.2.3", "12345678901234567890123", "1e", "abc"])
def test_unusual_numbers_ask_for_the_scanf_path(thsp, cuda, tmp_path, bad):
    from arm_spmv_b200 import host as H
    ent = [(1, 1, "1.5")] * 300 + [(2, 2, bad)] + [(1, 2, "2")] * 300
    path = str(tmp_path / "m.mtx")
    write_mtx(path, 2, 2, ent)
    with pytest.raises(H.MatrixMarketNeedsScanf):
        H.COOMatrixRead(path)


def test_short_file_asks_for_the_scanf_path(thsp, cuda, tmp_path):
    from arm_spmv_b200 import host as H
    path = str(tmp_path / "m.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n3 3 4\n1 1 1.0\n2 2 2.0\n3 3\n")
    with pytest.raises(H.MatrixMarketNeedsScanf):
        H.COOMatrixRead(path)


@pytest.mark.parametrize("case", ["api_lap5_12", "api_rand90"])
def test_golden_file_read_by_the_reference(thsp, cuda, case):
    """tests/golden/<case>/matrix.mtx was read and converted by the unmodified reference (make_api_golden.py); the
    CSR arrays it produced are the fixture.  GPU parser + GPU COO->CSR must reproduce them bit for bit."""
    from arm_spmv_b200 import host as H
    gold = os.path.join(ROOT, "tests", "golden", case)
    A = H.COOMatrixRead(os.path.join(gold, "matrix.mtx"))
    B = H.CSRMatrix(A)
    assert np.array_equal(B.row_ptr.cpu().numpy(), np.fromfile(os.path.join(gold, "csr_row_ptr.i32"), np.int32))
    assert np.array_equal(B.col_ind.cpu().numpy(), np.fromfile(os.path.join(gold, "csr_col_ind.i32"), np.int32))
    assert np.array_equal(bits(B.values.cpu().numpy()), bits(np.fromfile(os.path.join(gold, "csr_values.f64"), np.float64)))
    # and the text as Python reads it
    rows, cols, nz, body = H.mtx_split(os.path.join(gold, "matrix.mtx"))
    tok = body.split()
    assert A.nnz == nz and len(tok) >= 3 * nz
    assert np.array_equal(A.row_ind.cpu().numpy(), np.array([int(t) - 1 for t in tok[0:3 * nz:3]], np.int32))
    assert np.array_equal(A.col_ind.cpu().numpy(), np.array([int(t) - 1 for t in tok[1:3 * nz:3]], np.int32))
    assert np.array_equal(bits(A.values.cpu().numpy()), bits([float(t) for t in tok[2:3 * nz:3]]))
