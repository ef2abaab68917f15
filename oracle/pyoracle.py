"""numpy front-ends for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -> oracle/liboracle.so   the C restatement (oracle.c), prefix ``oracle_``
* ``Ref``     -> oracle/_ref/libref.so the unmodified reference behind ref_shim.cpp, prefix ``ref_``

Both expose the same method names so a test can run either.  Only tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _Base:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = C.CDLL(self.path)

    def fn(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    # ---- SpMV (all: y += A x on a copy of y, returned)
    def coo_spmv(self, nrow, ncol, ri, ci, v, x, y):
        y = _d(y).copy()
        ri, ci, v, x = _i(ri), _i(ci), _d(v), _d(x)
        if self.prefix == "ref_":
            self.fn("coo_spmv")(C.c_int(nrow), C.c_int(ncol), C.c_int(len(v)), _vp(ri), _vp(ci), _vp(v), _vp(x), _vp(y))
        else:
            self.fn("coo_spmv")(C.c_int(len(v)), _vp(ri), _vp(ci), _vp(v), _vp(x), _vp(y))
        return y

    def csr_spmv(self, nrow, ncol, rp, ci, v, x, y):
        y = _d(y).copy()
        rp, ci, v, x = _i(rp), _i(ci), _d(v), _d(x)
        if self.prefix == "ref_":
            self.fn("csr_spmv")(C.c_int(nrow), C.c_int(ncol), _vp(rp), _vp(ci), _vp(v), _vp(x), _vp(y))
        else:
            self.fn("csr_spmv")(C.c_int(nrow), _vp(rp), _vp(ci), _vp(v), _vp(x), _vp(y))
        return y

    def csc_spmv(self, nrow, ncol, cp, ri, v, x, y):
        y = _d(y).copy()
        cp, ri, v, x = _i(cp), _i(ri), _d(v), _d(x)
        if self.prefix == "ref_":
            self.fn("csc_spmv")(C.c_int(nrow), C.c_int(ncol), _vp(cp), _vp(ri), _vp(v), _vp(x), _vp(y))
        else:
            self.fn("csc_spmv")(C.c_int(ncol), _vp(cp), _vp(ri), _vp(v), _vp(x), _vp(y))
        return y

    def ell_spmv(self, nrow, ncol, width, ci, v, x, y):
        y = _d(y).copy()
        ci, v, x = _i(ci), _d(v), _d(x)
        if self.prefix == "ref_":
            self.fn("ell_spmv")(C.c_int(nrow), C.c_int(ncol), C.c_int(width), _vp(ci), _vp(v), _vp(x), _vp(y))
        else:
            self.fn("ell_spmv")(C.c_int(nrow), C.c_int(width), _vp(ci), _vp(v), _vp(x), _vp(y))
        return y

    def dia_spmv(self, nrow, ncol, off, v, x, y):
        y = _d(y).copy()
        off, v, x = _i(off), _d(v), _d(x)
        if self.prefix == "ref_":
            self.fn("dia_spmv")(C.c_int(nrow), C.c_int(ncol), C.c_int(len(off)), _vp(off), _vp(v), _vp(x), _vp(y))
        else:
            self.fn("dia_spmv")(C.c_int(nrow), C.c_int(len(off)), _vp(off), _vp(v), _vp(x), _vp(y))
        return y

    # ---- conversions
    def coo2csr(self, nrow, ncol, ri, ci, v):
        ri, ci, v = _i(ri), _i(ci), _d(v)
        nnz = len(v)
        rp = np.zeros(nrow + 1, np.int32)
        co = np.zeros(nnz, np.int32)
        va = np.zeros(nnz, np.float64)
        dg = np.zeros(max(nrow, 1), np.float64)
        nd = self.fn("coo2csr", C.c_int)(C.c_int(nrow), C.c_int(ncol), C.c_int(nnz), _vp(ri), _vp(ci), _vp(v),
                                         _vp(rp), _vp(co), _vp(va), _vp(dg))
        if nd < 0:
            raise ValueError("reference overruns diagonal[] on this input")
        return rp, co, va, dg[:min(nd, nrow)].copy()

    def coo2csc(self, nrow, ncol, ri, ci, v):
        ri, ci, v = _i(ri), _i(ci), _d(v)
        nnz = len(v)
        cp = np.zeros(ncol + 1, np.int32)
        ro = np.zeros(nnz, np.int32)
        va = np.zeros(nnz, np.float64)
        self.fn("coo2csc")(C.c_int(nrow), C.c_int(ncol), C.c_int(nnz), _vp(ri), _vp(ci), _vp(v), _vp(cp), _vp(ro), _vp(va))
        return cp, ro, va

    def coo2ell(self, nrow, ncol, ri, ci, v):
        ri, ci, v = _i(ri), _i(ci), _d(v)
        nnz = len(v)
        dg = np.zeros(max(nrow, 1), np.float64)
        if self.prefix == "ref_":
            f = self.fn("coo2ell", C.c_int)
            k = f(C.c_int(nrow), C.c_int(ncol), C.c_int(nnz), _vp(ri), _vp(ci), _vp(v), C.c_int(0), None, None, None)
            if k < 0:
                raise ValueError("reference overruns diagonal[] on this input")
            co = np.zeros(nrow * k, np.int32)
            va = np.zeros(nrow * k, np.float64)
            f(C.c_int(nrow), C.c_int(ncol), C.c_int(nnz), _vp(ri), _vp(ci), _vp(v), C.c_int(nrow * k), _vp(co), _vp(va), _vp(dg))
            nd = int(np.sum(ri == ci))
        else:
            k = self.fn("coo2ell_width", C.c_int)(C.c_int(nrow), C.c_int(nnz), _vp(ri))
            co = np.zeros(nrow * k, np.int32)
            va = np.zeros(nrow * k, np.float64)
            nd = self.fn("coo2ell", C.c_int)(C.c_int(nrow), C.c_int(ncol), C.c_int(nnz), _vp(ri), _vp(ci), _vp(v),
                                             C.c_int(k), _vp(co), _vp(va), _vp(dg))
        return k, co, va, dg[:min(nd, nrow)].copy()

    def csr2dia(self, nrow, ncol, rp, ci, v):
        rp, ci, v = _i(rp), _i(ci), _d(v)
        if self.prefix == "ref_":
            f = self.fn("csr2dia", C.c_int)
            nd = f(C.c_int(nrow), C.c_int(ncol), _vp(rp), _vp(ci), _vp(v), C.c_int(0), None, None)
            if nd < 0:
                raise ValueError("reference writes out of bounds on a (0, ncol-1) entry")
            off = np.zeros(nd, np.int32)
            va = np.zeros(nd * nrow, np.float64)
            f(C.c_int(nrow), C.c_int(ncol), _vp(rp), _vp(ci), _vp(v), C.c_int(nd), _vp(off), _vp(va))
        else:
            nd = self.fn("csr2dia_offsets", C.c_int)(C.c_int(nrow), C.c_int(ncol), _vp(rp), _vp(ci), None)
            off = np.zeros(nd, np.int32)
            va = np.zeros(nd * nrow, np.float64)
            self.fn("csr2dia_offsets", C.c_int)(C.c_int(nrow), C.c_int(ncol), _vp(rp), _vp(ci), _vp(off))
            self.fn("csr2dia_fill")(C.c_int(nrow), C.c_int(ncol), _vp(rp), _vp(ci), _vp(v), C.c_int(nd), _vp(off), _vp(va))
        return off, va

    # ---- vectors
    def dot(self, x, y):
        x, y = _d(x), _d(y)
        return float(self.fn("dot", C.c_double)(C.c_int(len(x)), _vp(x), _vp(y)))

    def axpby(self, alpha, x, beta, y):
        x, y = _d(x), _d(y)
        w = np.zeros_like(x)
        self.fn("axpby")(C.c_int(len(x)), C.c_double(alpha), _vp(x), C.c_double(beta), _vp(y), _vp(w))
        return w

    def fill(self, n, a):
        v = np.empty(n, np.float64)
        self.fn("fill")(C.c_int(n), C.c_double(a), _vp(v))
        return v

    def scale(self, a, v):
        v = _d(v).copy()
        self.fn("scale")(C.c_int(len(v)), C.c_double(a), _vp(v))
        return v

    def shift(self, a, v):
        v = _d(v).copy()
        self.fn("shift")(C.c_int(len(v)), C.c_double(a), _vp(v))
        return v

    def add_scaled(self, a, x, v):
        v, x = _d(v).copy(), _d(x)
        self.fn("add_scaled")(C.c_int(len(v)), C.c_double(a), _vp(x), _vp(v))
        return v

    def add2_scaled(self, a, x, b, y, v):
        v, x, y = _d(v).copy(), _d(x), _d(y)
        self.fn("add2_scaled")(C.c_int(len(v)), C.c_double(a), _vp(x), C.c_double(b), _vp(y), _vp(v))
        return v

    def check_vector(self, x, y):
        x, y = _d(x), _d(y)
        return bool(self.fn("check_vector", C.c_int)(C.c_int(len(x)), _vp(x), C.c_int(len(y)), _vp(y)))


class Oracle(_Base):
    prefix = "oracle_"
    path = os.path.join(_HERE, "liboracle.so")

    def csr_spmv_f32(self, nrow, rp, ci, v, x, y):
        y = np.ascontiguousarray(y, np.float32).copy()
        rp, ci = _i(rp), _i(ci)
        v = np.ascontiguousarray(v, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        self.fn("csr_spmv_f32")(C.c_int(nrow), _vp(rp), _vp(ci), _vp(v), _vp(x), _vp(y))
        return y

    def ell_spmv_f32(self, nrow, width, ci, v, x, y):
        y = np.ascontiguousarray(y, np.float32).copy()
        ci = _i(ci)
        v = np.ascontiguousarray(v, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        self.fn("ell_spmv_f32")(C.c_int(nrow), C.c_int(width), _vp(ci), _vp(v), _vp(x), _vp(y))
        return y

    def tile_sumsq(self, y):
        y = _d(y)
        out = np.zeros((len(y) + 31) // 32, np.float64)
        self.fn("tile_sumsq")(C.c_int64(len(y)), _vp(y), _vp(out))
        return out

    def tree_sum(self, vals):
        vals = _d(vals)
        return float(self.fn("tree_sum", C.c_double)(C.c_int64(len(vals)), _vp(vals)))

    def hash_f64(self, v, first=0):
        v = _d(v)
        return int(self.fn("hash_f64", C.c_uint64)(C.c_int64(len(v)), _vp(v), C.c_uint64(first)))

    def symgs(self, color_ptr, perm, rp, ci, va, diag, r, x):
        x = _d(x).copy()
        cp, perm, rp, ci = _i(color_ptr), _i(perm), _i(rp), _i(ci)
        va, diag, r = _d(va), _d(diag), _d(r)
        self.fn("symgs")(C.c_int(len(cp) - 1), _vp(cp), _vp(perm), _vp(rp), _vp(ci), _vp(va), _vp(diag), _vp(r), _vp(x))
        return x

    def symgs_sequential(self, rp, ci, va, diag, r, x):
        x = _d(x).copy()
        rp, ci, va, diag, r = _i(rp), _i(ci), _d(va), _d(diag), _d(r)
        self.fn("symgs_sequential")(C.c_int(len(rp) - 1), _vp(rp), _vp(ci), _vp(va), _vp(diag), _vp(r), _vp(x))
        return x

    def dot_canonical(self, a, b):
        a, b = _d(a), _d(b)
        return float(self.fn("dot_canonical", C.c_double)(C.c_int64(len(a)), _vp(a), _vp(b)))

    def cg(self, rp, ci, va, diag, b, x0, maxit, tol, precond=0, color_ptr=None, perm=None):
        rp, ci, va, b = _i(rp), _i(ci), _d(va), _d(b)
        x = _d(x0).copy()
        diag = None if diag is None else _d(diag)
        cp = None if color_ptr is None else _i(color_ptr)
        pm = None if perm is None else _i(perm)
        rel = C.c_double(0.0)
        it = self.fn("cg", C.c_int)(C.c_int(len(rp) - 1), _vp(rp), _vp(ci), _vp(va), _vp(diag), C.c_int(precond),
                                    C.c_int(0 if cp is None else len(cp) - 1), _vp(cp), _vp(pm), _vp(b), _vp(x), C.c_int(maxit),
                                    C.c_double(tol), C.byref(rel))
        return x, int(it), rel.value

    def partition(self, n, nparts, part):
        s, c = C.c_int(), C.c_int()
        self.fn("partition")(C.c_int(n), C.c_int(nparts), C.c_int(part), C.byref(s), C.byref(c))
        return s.value, c.value

    def csr_slice(self, rp, start, count):
        rp = _i(rp)
        sub = np.zeros(count + 1, np.int32)
        nnz = self.fn("csr_slice", C.c_int)(_vp(rp), C.c_int(start), C.c_int(count), _vp(sub))
        return sub, nnz

    def gen_vector(self, n, seed):
        v = np.empty(n, np.float64)
        self.fn("gen_vector")(C.c_int64(n), C.c_uint64(seed), _vp(v))
        return v

    def gen_stencil27_csr(self, n, r0=0, r1=None):
        r1 = n ** 3 if r1 is None else r1
        rp = np.zeros(r1 - r0 + 1, np.int32)
        f = self.fn("gen_stencil27_csr", C.c_int64)
        nnz = f(C.c_int(n), C.c_int64(r0), C.c_int64(r1), _vp(rp), None, None)
        ci = np.zeros(nnz, np.int32)
        va = np.zeros(nnz, np.float64)
        f(C.c_int(n), C.c_int64(r0), C.c_int64(r1), _vp(rp), _vp(ci), _vp(va))
        return rp, ci, va

    def gen_lap5_coo(self, n):
        f = self.fn("gen_lap5_coo", C.c_int)
        nnz = f(C.c_int(n), None, None, None)
        ri = np.zeros(nnz, np.int32)
        ci = np.zeros(nnz, np.int32)
        va = np.zeros(nnz, np.float64)
        f(C.c_int(n), _vp(ri), _vp(ci), _vp(va))
        return ri, ci, va

    def gen_uniform_coo(self, nrow, ncol, nnz, seed):
        ri = np.zeros(nnz, np.int32)
        ci = np.zeros(nnz, np.int32)
        va = np.zeros(nnz, np.float64)
        self.fn("gen_uniform_coo")(C.c_int(nrow), C.c_int(ncol), C.c_int64(nnz), C.c_uint64(seed), _vp(ri), _vp(ci), _vp(va))
        return ri, ci, va

    def gen_rmat_coo(self, scale, nnz, seed):
        ri = np.zeros(nnz, np.int32)
        ci = np.zeros(nnz, np.int32)
        va = np.zeros(nnz, np.float64)
        self.fn("gen_rmat_coo")(C.c_int(scale), C.c_int64(nnz), C.c_uint64(seed), _vp(ri), _vp(ci), _vp(va))
        return ri, ci, va


class Ref(_Base):
    prefix = "ref_"
    path = os.path.join(_HERE, "_ref", "libref.so")

    def set_threads(self, n):
        self.fn("set_threads")(C.c_int(n))

    def max_threads(self):
        return int(self.fn("max_threads", C.c_int)())

    def coo_read(self, path):
        nr, nc, nz = C.c_int(), C.c_int(), C.c_int()
        f = self.fn("coo_read", C.c_int)
        f(path.encode(), C.byref(nr), C.byref(nc), C.byref(nz), C.c_int(0), None, None, None)
        ri = np.zeros(nz.value, np.int32)
        ci = np.zeros(nz.value, np.int32)
        va = np.zeros(nz.value, np.float64)
        f(path.encode(), C.byref(nr), C.byref(nc), C.byref(nz), C.c_int(nz.value), _vp(ri), _vp(ci), _vp(va))
        return nr.value, nc.value, ri, ci, va

    def time_csr_spmv(self, nrow, ncol, rp, ci, v, x, reps):
        rp, ci, v, x = _i(rp), _i(ci), _d(v), _d(x)
        y = np.zeros(nrow, np.float64)
        return float(self.fn("time_csr_spmv", C.c_double)(C.c_int(nrow), C.c_int(ncol), _vp(rp), _vp(ci), _vp(v), _vp(x),
                                                           _vp(y), C.c_int(reps)))

    def time_ell_spmv(self, nrow, ncol, width, ci, v, x, reps):
        ci, v, x = _i(ci), _d(v), _d(x)
        y = np.zeros(nrow, np.float64)
        return float(self.fn("time_ell_spmv", C.c_double)(C.c_int(nrow), C.c_int(ncol), C.c_int(width), _vp(ci), _vp(v),
                                                           _vp(x), _vp(y), C.c_int(reps)))

    def time_power_iteration(self, n, rp, ci, v, x, steps):
        rp, ci, v = _i(rp), _i(ci), _d(v)
        x = _d(x).copy()
        y = np.zeros(n, np.float64)
        nrm = C.c_double()
        dt = float(self.fn("time_power_iteration", C.c_double)(C.c_int(n), _vp(rp), _vp(ci), _vp(v), _vp(x), _vp(y),
                                                                C.c_int(steps), C.byref(nrm)))
        return dt, nrm.value, x


class RefO3(Ref):
    """The same reference sources built with -O3 -march=x86-64-v3 (oracle/Makefile): a timing baseline only -
    FMA contraction changes last bits, so parity is never checked against it."""
    path = os.path.join(_HERE, "_ref", "libref_o3.so")
    FLAGS = "-O3 -march=x86-64-v3 -fopenmp -DUSE_OPENMP"

    @staticmethod
    def runnable():
        """The host CPU has what x86-64-v3 code needs (the library was built on another machine)."""
        if not os.path.exists(RefO3.path):
            return False
        try:
            flags = set()
            for ln in open("/proc/cpuinfo"):
                if ln.startswith("flags"):
                    flags = set(ln.split(":", 1)[1].split())
                    break
            return {"avx2", "fma", "bmi2", "f16c", "movbe"} <= flags
        except OSError:
            return False
