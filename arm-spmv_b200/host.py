"""Python mirror of the reference's operator interface, on torch CUDA tensors.

Same names, argument order and semantics as the reference's headers (include/matrix.h,
vector.h, mat_vec.h, vec_vec.h): ``y += A x``, converting constructors, ``vec_dot`` /
``vec_axpby``.  Every method is a thin call into the C ABI (lib.py -> libthsparse_cuda.so) on
torch's current stream; torch only provides device memory and streams.  Nothing here computes
on the CPU.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L
from .lib import check, current_stream, load, ptr

I32 = torch.int32
F64 = torch.float64


def _dev(device=None):
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _as(t, dtype, device=None):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=_dev(device), dtype=dtype).contiguous()


class Vector:
    """include/vector.h:4-26."""

    def __init__(self, n_or_values=0, device=None):
        if isinstance(n_or_values, int):
            self.values = torch.empty(n_or_values, dtype=F64, device=_dev(device))
        else:
            self.values = _as(n_or_values, F64, device)

    @property
    def size(self):
        return self.values.numel()

    def Resize(self, n):  # src/vector.cpp:51-57 (contents are not preserved)
        self.values = torch.empty(n, dtype=F64, device=self.values.device)

    def Fill(self, a):
        check(load().thsp_fill_f64(C.c_int64(self.size), C.c_double(a), ptr(self.values), current_stream()))

    def Scale(self, a):
        check(load().thsp_scale_f64(C.c_int64(self.size), C.c_double(a), ptr(self.values), current_stream()))

    def Shift(self, a):
        check(load().thsp_shift_f64(C.c_int64(self.size), C.c_double(a), ptr(self.values), current_stream()))

    def Copy(self, x: "Vector"):
        check(load().thsp_copy_f64(C.c_int64(self.size), ptr(x.values), ptr(self.values), current_stream()))

    def AddScaled(self, a, x: "Vector"):
        check(load().thsp_add_scaled_f64(C.c_int64(self.size), C.c_double(a), ptr(x.values), ptr(self.values), current_stream()))

    def Add2Scaled(self, a, x: "Vector", b, y: "Vector"):
        check(load().thsp_add2_scaled_f64(C.c_int64(self.size), C.c_double(a), ptr(x.values), C.c_double(b), ptr(y.values),
                                          ptr(self.values), current_stream()))

    def FillRandom(self, seed=1):
        """The reference draws glibc rand() on the host (src/vector.cpp:65-69); here a counter hash
        generates the same kind of U[0,1) vector on the device (oracle_gen_vector is its CPU twin)."""
        check(load().thsp_gen_vector_f64(C.c_int64(self.size), C.c_uint64(seed), ptr(self.values), current_stream()))


def checkVector(x: Vector, y: Vector) -> bool:  # src/vector.cpp:161-171
    ok = C.c_int(0)
    check(load().thsp_check_vector_f64(C.c_int64(x.size), ptr(x.values), C.c_int64(y.size), ptr(y.values), C.byref(ok),
                                       current_stream()))
    return bool(ok.value)


def vec_dot(x: Vector, y: Vector) -> float:  # src/vec_vec.cpp:15-29 (n = x.size)
    out = C.c_double(0.0)
    check(load().thsp_dot_f64(C.c_int64(x.size), ptr(x.values), ptr(y.values), C.byref(out), current_stream()))
    return out.value


def vec_axpby(alpha, x: Vector, beta, y: Vector, w: Vector) -> None:  # src/vec_vec.cpp:31-94 (n = w.size)
    check(load().thsp_axpby_f64(C.c_int64(w.size), C.c_double(alpha), ptr(x.values), C.c_double(beta), ptr(y.values),
                                ptr(w.values), current_stream()))


class COOMatrix:
    """include/matrix.h:7-26."""

    def __init__(self, nrow=0, ncol=0, row_ind=None, col_ind=None, values=None, device=None):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.row_ind = _as(row_ind if row_ind is not None else [], I32, device)
        self.col_ind = _as(col_ind if col_ind is not None else [], I32, device)
        self.values = _as(values if values is not None else [], F64, device)

    @property
    def nnz(self):
        return self.values.numel()


class CSRMatrix:
    """include/matrix.h:28-48; CSRMatrix(COOMatrix) = src/matrix.cpp:115-154."""

    def __init__(self, A=None, *, nrow=0, ncol=0, row_ptr=None, col_ind=None, values=None, diagonal=None, device=None):
        self._plan = None
        if isinstance(A, COOMatrix):
            dev = A.values.device
            self.nrow, self.ncol = A.nrow, A.ncol
            self.row_ptr = torch.empty(A.nrow + 1, dtype=I32, device=dev)
            self.col_ind = torch.empty(A.nnz, dtype=I32, device=dev)
            self.values = torch.empty(A.nnz, dtype=F64, device=dev)
            self.diagonal = torch.empty(max(A.nrow, 1), dtype=F64, device=dev)
            nd = C.c_int(0)
            check(load().thsp_coo2csr(A.nrow, A.ncol, A.nnz, ptr(A.row_ind), ptr(A.col_ind), ptr(A.values), ptr(self.row_ptr),
                                      ptr(self.col_ind), ptr(self.values), ptr(self.diagonal), C.byref(nd), current_stream()))
            self.ndiag = nd.value
        else:
            self.nrow, self.ncol = int(nrow), int(ncol)
            self.row_ptr = _as(row_ptr, I32, device)
            self.col_ind = _as(col_ind, I32, device)
            self.values = values if torch.is_tensor(values) and values.is_cuda else _as(values, F64, device)
            self.diagonal = diagonal
            self.ndiag = 0

    @property
    def nnz(self):
        return self.col_ind.numel()

    def plan(self, autotune: bool = False):
        """Row-length histogram -> kernel choice (or a measurement when autotune=True or
        THSP_AUTOTUNE=1); cached until the arrays are replaced."""
        key = (self.row_ptr.data_ptr(), self.col_ind.data_ptr(), self.values.data_ptr(), self.nrow, self.nnz)
        if self._plan is None or self._plan[0] != key:
            self.free_plan()
            h = C.c_void_p()
            check(load().thsp_csr_plan_create(C.byref(h), self.nrow, self.ncol, self.nnz, ptr(self.row_ptr), ptr(self.col_ind),
                                              ptr(self.values), self.values.element_size(), current_stream()))
            self._plan = (key, h)
            import os
            if autotune or os.environ.get("THSP_AUTOTUNE") == "1":
                check(load().thsp_csr_plan_autotune(h, current_stream()))
        return self._plan[1]

    def plan_kernel(self):
        k, l = C.c_int(), C.c_int()
        check(load().thsp_csr_plan_kernel(self.plan(), C.byref(k), C.byref(l)))
        return L.KERNEL_NAMES[k.value], l.value

    def free_plan(self):
        if self._plan is not None:
            load().thsp_csr_plan_destroy(self._plan[1])
            self._plan = None

    def __del__(self):
        try:
            self.free_plan()
        except Exception:
            pass


class CSCMatrix:
    """include/matrix.h:50-69; CSCMatrix(COOMatrix) = src/matrix.cpp:295-325."""

    def __init__(self, A=None, *, nrow=0, ncol=0, col_ptr=None, row_ind=None, values=None, device=None):
        if isinstance(A, COOMatrix):
            dev = A.values.device
            self.nrow, self.ncol = A.nrow, A.ncol
            self.col_ptr = torch.empty(A.ncol + 1, dtype=I32, device=dev)
            self.row_ind = torch.empty(A.nnz, dtype=I32, device=dev)
            self.values = torch.empty(A.nnz, dtype=F64, device=dev)
            check(load().thsp_coo2csc(A.nrow, A.ncol, A.nnz, ptr(A.row_ind), ptr(A.col_ind), ptr(A.values), ptr(self.col_ptr),
                                      ptr(self.row_ind), ptr(self.values), current_stream()))
        else:
            self.nrow, self.ncol = int(nrow), int(ncol)
            self.col_ptr = _as(col_ptr, I32, device)
            self.row_ind = _as(row_ind, I32, device)
            self.values = _as(values, F64, device)

    @property
    def nnz(self):
        return self.row_ind.numel()


class ELLMatrix:
    """include/matrix.h:71-93 (column-major slab); ELLMatrix(COOMatrix) = src/matrix.cpp:450-500."""

    def __init__(self, A=None, *, nrow=0, ncol=0, nnz=0, nonzeros_in_row=0, col_ind=None, values=None, diagonal=None,
                 device=None):
        if isinstance(A, COOMatrix):
            dev = A.values.device
            self.nrow, self.ncol, self.nnz = A.nrow, A.ncol, A.nnz
            k = C.c_int(0)
            check(load().thsp_coo2ell_prepare(A.nrow, A.ncol, A.nnz, ptr(A.row_ind), ptr(A.col_ind), ptr(A.values), C.byref(k),
                                              current_stream()))
            self.nonzeros_in_row = k.value
            self.col_ind = torch.empty(A.nrow * k.value, dtype=I32, device=dev)
            self.values = torch.empty(A.nrow * k.value, dtype=F64, device=dev)
            self.diagonal = torch.empty(max(A.nrow, 1), dtype=F64, device=dev)
            nd = C.c_int(0)
            check(load().thsp_coo2ell(A.nrow, A.ncol, A.nnz, ptr(A.row_ind), ptr(A.col_ind), ptr(A.values), k.value,
                                      ptr(self.col_ind), ptr(self.values), ptr(self.diagonal), C.byref(nd), current_stream()))
            self.ndiag = nd.value
        else:
            self.nrow, self.ncol, self.nnz, self.nonzeros_in_row = int(nrow), int(ncol), int(nnz), int(nonzeros_in_row)
            self.col_ind = _as(col_ind, I32, device)
            self.values = _as(values, F64, device)
            self.diagonal = diagonal
            self.ndiag = 0


class DIAMatrix:
    """include/matrix.h:117-137 (row-major diagonals); DIAMatrix(CSRMatrix) = src/matrix.cpp:673-726."""

    def __init__(self, A=None, *, nrow=0, ncol=0, offsets=None, values=None, device=None):
        if isinstance(A, CSRMatrix):
            dev = A.values.device
            self.nrow, self.ncol, self.nnz = A.nrow, A.ncol, A.nnz
            nd = C.c_int(0)
            check(load().thsp_csr2dia_offsets(A.nrow, A.ncol, ptr(A.row_ptr), ptr(A.col_ind), C.byref(nd), None, 0,
                                              current_stream()))
            self.ndiags = nd.value
            self.offsets = torch.empty(max(nd.value, 1), dtype=I32, device=dev)[:nd.value]
            self.values = torch.empty(A.nrow * nd.value, dtype=F64, device=dev)
            check(load().thsp_csr2dia_offsets(A.nrow, A.ncol, ptr(A.row_ptr), ptr(A.col_ind), C.byref(nd), ptr(self.offsets),
                                              nd.value, current_stream()))
            check(load().thsp_csr2dia_fill(A.nrow, A.ncol, ptr(A.row_ptr), ptr(A.col_ind), ptr(A.values), nd.value,
                                           ptr(self.offsets), ptr(self.values), current_stream()))
        else:
            self.nrow, self.ncol = int(nrow), int(ncol)
            self.offsets = _as(offsets, I32, device)
            self.values = _as(values, F64, device)
            self.ndiags = self.offsets.numel()
            self.nnz = 0


# ---- include/mat_vec.h:7-11 : y += A x ------------------------------------------------------
def COOMatirxMatVector(A: COOMatrix, x: Vector, y: Vector) -> None:  # [sic] the reference's spelling
    check(load().thsp_coo_spmv_f64(A.nrow, A.ncol, A.nnz, ptr(A.row_ind), ptr(A.col_ind), ptr(A.values), ptr(x.values),
                                   ptr(y.values), current_stream()))


def CSRMatrixMatVector(A: CSRMatrix, x: Vector, y: Vector, accumulate: bool = True) -> None:
    fn = load().thsp_csr_plan_spmv_f64 if A.values.dtype == F64 else load().thsp_csr_plan_spmv_f32
    check(fn(A.plan(), ptr(x.values), ptr(y.values), 1 if accumulate else 0, current_stream()))


def CSCMatrixMatVector(A: CSCMatrix, x: Vector, y: Vector) -> None:
    check(load().thsp_csc_spmv_f64(A.nrow, A.ncol, A.nnz, ptr(A.col_ptr), ptr(A.row_ind), ptr(A.values), ptr(x.values),
                                   ptr(y.values), current_stream()))


def ELLMatrixMatVector(A: ELLMatrix, x: Vector, y: Vector) -> None:
    check(load().thsp_ell_spmv_f64(A.nrow, A.ncol, A.nonzeros_in_row, ptr(A.col_ind), ptr(A.values), ptr(x.values),
                                   ptr(y.values), current_stream()))


def DIAMatrixMatVector(A: DIAMatrix, x: Vector, y: Vector) -> None:
    check(load().thsp_dia_spmv_f64(A.nrow, A.ncol, A.ndiags, ptr(A.offsets), ptr(A.values), ptr(x.values), ptr(y.values),
                                   current_stream()))


def csr_spmv_kernel(kernel: int, lanes: int, A: CSRMatrix, x: torch.Tensor, y: torch.Tensor, accumulate: bool = True) -> None:
    """Forced-kernel CSR SpMV (tests / tuning)."""
    fn = load().thsp_csr_spmv_kernel_f64 if A.values.dtype == F64 else load().thsp_csr_spmv_kernel_f32
    check(fn(kernel, lanes, A.nrow, A.ncol, A.nnz, ptr(A.row_ptr), ptr(A.col_ind), ptr(A.values), ptr(x), ptr(y),
             1 if accumulate else 0, current_stream()))


# ---- include/data_io.h:12 : Matrix Market reader ----------------------------------------------
class MatrixMarketNeedsScanf(ValueError):
    """The entry section holds something the GPU parser's fast conversions do not cover."""


def mtx_split(path: str):
    """Banner, comments and size line of a coordinate file (what mmio does on the CPU in the reference,
    src/mmio.cpp:109-229) -> (nrow, ncol, nnz, bytes of the entry section)."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    first = True
    while True:
        end = data.find(b"\n", pos)
        line = data[pos:] if end < 0 else data[pos:end]
        nxt = len(data) if end < 0 else end + 1
        if first:
            if not line.lower().startswith(b"%%matrixmarket"):
                raise ValueError("*** Could not process Matrix Market banner ***")
            words = line.lower().split()
            if len(words) >= 4 and words[1] == b"matrix" and words[2] == b"coordinate" and words[3] == b"complex":
                raise ValueError("Sorry, this application does not support Market Market type: [" + b" ".join(words[1:]).decode() + "]")
            first = False
        elif line.strip() and not line.startswith(b"%"):
            rows, cols, nz = (int(v) for v in line.split()[:3])
            return rows, cols, nz, data[nxt:]
        if end < 0:
            raise ValueError("no size line")
        pos = nxt


def COOMatrixRead(path: str, device=None) -> COOMatrix:
    """COOMatrixRead (src/data_io.cpp:45-105) with the entry loop (:83-88) on the GPU (thsp_mtx_parse_coo)."""
    rows, cols, nz, body = mtx_split(path)
    dev = _dev(device)
    check(load().thsp_prepare_conversions(rows, cols, nz, current_stream()))   # as the C++ reader does (csrc/host/data_io.cpp)
    ri = torch.empty(nz, dtype=I32, device=dev)
    ci = torch.empty(nz, dtype=I32, device=dev)
    va = torch.empty(nz, dtype=F64, device=dev)
    status = C.c_int(1)
    check(load().thsp_mtx_parse_coo(C.c_char_p(body), C.c_size_t(len(body)), nz, ptr(ri), ptr(ci), ptr(va), C.byref(status), current_stream()))
    if status.value != 0:
        raise MatrixMarketNeedsScanf(path)
    return COOMatrix(rows, cols, ri, ci, va)


# ---- synthetic inputs ------------------------------------------------------
def stencil27_csr(n: int, row_begin: int = 0, row_end: int | None = None, device=None) -> CSRMatrix:
    N = n ** 3
    row_end = N if row_end is None else row_end
    dev = _dev(device)
    nnz = int(load().thsp_stencil27_nnz(n, row_begin, row_end))
    rp = torch.empty(row_end - row_begin + 1, dtype=I32, device=dev)
    ci = torch.empty(nnz, dtype=I32, device=dev)
    va = torch.empty(nnz, dtype=F64, device=dev)
    check(load().thsp_gen_stencil27_csr(n, C.c_int64(row_begin), C.c_int64(row_end), ptr(rp), ptr(ci), ptr(va), current_stream()))
    return CSRMatrix(nrow=row_end - row_begin, ncol=N, row_ptr=rp, col_ind=ci, values=va)


def stencil27_ell(n: int, device=None) -> ELLMatrix:
    N = n ** 3
    dev = _dev(device)
    ci = torch.empty(N * 27, dtype=I32, device=dev)
    va = torch.empty(N * 27, dtype=F64, device=dev)
    check(load().thsp_gen_stencil27_ell(n, ptr(ci), ptr(va), current_stream()))
    return ELLMatrix(nrow=N, ncol=N, nnz=int(load().thsp_stencil27_nnz(n, 0, N)), nonzeros_in_row=27, col_ind=ci, values=va)


def stencil27_coo(n: int, device=None) -> COOMatrix:
    N = n ** 3
    dev = _dev(device)
    nnz = int(load().thsp_stencil27_nnz(n, 0, N))
    ri = torch.empty(nnz, dtype=I32, device=dev)
    ci = torch.empty(nnz, dtype=I32, device=dev)
    va = torch.empty(nnz, dtype=F64, device=dev)
    check(load().thsp_gen_stencil27_coo(n, ptr(ri), ptr(ci), ptr(va), current_stream()))
    return COOMatrix(N, N, ri, ci, va)


def lap5_coo(n: int, device=None) -> COOMatrix:
    dev = _dev(device)
    nnz = int(load().thsp_lap5_nnz(n))
    ri = torch.empty(nnz, dtype=I32, device=dev)
    ci = torch.empty(nnz, dtype=I32, device=dev)
    va = torch.empty(nnz, dtype=F64, device=dev)
    check(load().thsp_gen_lap5_coo(n, ptr(ri), ptr(ci), ptr(va), current_stream()))
    return COOMatrix(n * n, n * n, ri, ci, va)


def uniform_coo(nrow: int, ncol: int, nnz: int, seed: int, device=None) -> COOMatrix:
    dev = _dev(device)
    ri = torch.empty(nnz, dtype=I32, device=dev)
    ci = torch.empty(nnz, dtype=I32, device=dev)
    va = torch.empty(nnz, dtype=F64, device=dev)
    check(load().thsp_gen_uniform_coo(nrow, ncol, C.c_int64(nnz), C.c_uint64(seed), ptr(ri), ptr(ci), ptr(va), current_stream()))
    return COOMatrix(nrow, ncol, ri, ci, va)


def rmat_coo(scale: int, nnz: int, seed: int, device=None) -> COOMatrix:
    dev = _dev(device)
    ri = torch.empty(nnz, dtype=I32, device=dev)
    ci = torch.empty(nnz, dtype=I32, device=dev)
    va = torch.empty(nnz, dtype=F64, device=dev)
    check(load().thsp_gen_rmat_coo(scale, C.c_int64(nnz), C.c_uint64(seed), ptr(ri), ptr(ci), ptr(va), current_stream()))
    return COOMatrix(1 << scale, 1 << scale, ri, ci, va)


def gen_vector(n: int, seed: int, device=None) -> Vector:
    v = Vector(n, device=device)
    v.FillRandom(seed)
    return v
