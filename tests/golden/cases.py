"""Seeded small inputs shared by the golden-fixture generator and the tests."""
import numpy as np


def _coo(nrow, ncol, nnz, seed, dup_frac=0.0, empty_rows=()):
    rs = np.random.RandomState(seed)
    ri = rs.randint(0, nrow, nnz).astype(np.int32)
    ci = rs.randint(0, ncol, nnz).astype(np.int32)
    if dup_frac > 0 and nnz > 4:
        k = int(nnz * dup_frac)
        src = rs.randint(0, nnz, k)
        dst = rs.randint(0, nnz, k)
        ri[dst], ci[dst] = ri[src], ci[src]
    for r in empty_rows:
        mask = ri == r
        ri[mask] = (r + 1) % nrow
    # keep the packed diagonal within the reference's nrow-sized buffer and avoid the
    # (0, ncol-1) corner entry on which the reference's DIA constructor writes out of bounds
    corner = (ri == 0) & (ci == ncol - 1)
    ci[corner] = 0
    va = rs.uniform(-1.0, 1.0, nnz)
    return ri, ci, va


def cases():
    out = {}
    out["kat4x5"] = dict(nrow=4, ncol=5, ri=np.array([3, 1, 0, 1, 3, 1, 0, 3], np.int32),
                         ci=np.array([4, 2, 0, 0, 3, 2, 3, 0], np.int32), va=np.arange(1.0, 9.0))
    ri, ci, va = _coo(37, 53, 400, 11, dup_frac=0.15, empty_rows=(5, 6, 36))
    out["rand37x53"] = dict(nrow=37, ncol=53, ri=ri, ci=ci, va=va)
    ri, ci, va = _coo(100, 7, 350, 12)
    out["tall100x7"] = dict(nrow=100, ncol=7, ri=ri, ci=ci, va=va)
    ri, ci, va = _coo(7, 100, 350, 13, dup_frac=0.05)
    out["wide7x100"] = dict(nrow=7, ncol=100, ri=ri, ci=ci, va=va)
    ri, ci, va = _coo(64, 64, 900, 14)
    ri[:600] = 17  # one long row, most others short or empty
    out["longrow64"] = dict(nrow=64, ncol=64, ri=ri, ci=ci, va=va)
    out["empty5x5"] = dict(nrow=5, ncol=5, ri=np.zeros(0, np.int32), ci=np.zeros(0, np.int32), va=np.zeros(0))
    ri, ci, va = _coo(1, 1, 1, 15)
    out["one1x1"] = dict(nrow=1, ncol=1, ri=ri, ci=ci, va=va, no_dia=True)  # (0,0) is the corner when ncol==1
    # 300 x 300 banded, sorted by row (the identity-permutation fast path), odd row lengths
    rs = np.random.RandomState(16)
    rows, cols = [], []
    for r in range(300):
        for d in (-17, -3, -1, 0, 1, 2, 40):
            c = r + d
            if 0 <= c < 300 and not (r == 0 and c == 299):
                rows.append(r)
                cols.append(c)
    out["band300"] = dict(nrow=300, ncol=300, ri=np.array(rows, np.int32), ci=np.array(cols, np.int32),
                          va=rs.uniform(-2, 2, len(rows)))
    for name, c in out.items():
        rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        c["x"] = rs.uniform(0.0, 1.0, c["ncol"])
        c["y0"] = rs.uniform(-1.0, 1.0, c["nrow"])
    return out


AXPBY_COEFFS = [(0.0, 2.5), (1.75, 0.0), (1.0, -0.3), (-1.0, 0.7), (0.4, 1.0), (-2.2, -1.0), (0.3, 0.9), (0.0, 0.0)]
ADD_SCALED_COEFFS = [0.0, 1.0, -1.0, 0.37]
ADD2_COEFFS = [(0.0, 0.5), (0.6, 0.0), (1.0, 0.25), (0.3, 1.0), (0.7, -1.3)]


def vec_inputs(n=1000, seed=99):
    rs = np.random.RandomState(seed)
    return rs.uniform(-1, 1, n), rs.uniform(-1, 1, n), rs.uniform(-1, 1, n)
