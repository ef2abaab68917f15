"""Helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np
import torch


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def assert_bits(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    if a.tobytes() != b.tobytes():
        bad = np.flatnonzero(a != b)
        raise AssertionError(f"{what}: {len(bad)} of {a.size} entries differ, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}")


def row_scale_csr(nrow, rp, ci, va, x):
    """sum_j |a_ij x_j| per row: the denominator of the per-row error."""
    p = np.abs(va * x[ci])
    s = np.add.reduceat(np.concatenate([p, [0.0]]), np.minimum(rp[:-1], len(p)))
    s[np.diff(rp) == 0] = 0.0
    return s


def row_scale_coo(nrow, ri, ci, va, x):
    return np.bincount(ri, weights=np.abs(va * x[ci]), minlength=nrow)


def max_row_error(y, y_ref, scale, y0=None):
    """max_i |y_i - yref_i| / (sum_j |a_ij x_j| + |y0_i|); rows with zero scale must match exactly."""
    den = scale + (np.abs(y0) if y0 is not None else 0.0)
    err = np.abs(y - y_ref)
    ok_zero = np.all(err[den == 0] == 0)
    assert ok_zero, "a row with no contributions changed"
    m = den > 0
    return float(np.max(err[m] / den[m])) if m.any() else 0.0
