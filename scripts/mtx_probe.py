"""BASELINE configs[0] ingest: 5-point Laplacian 1024^2 as a .mtx file (5.2 M entries, ~100 MB of text) read by the GPU
parser, and by the C++ drop-in reader with the GPU parser on / forced to the reference's scanf loop (bin/first_call is not
needed: bin/main prints its own progress; here only the reader is timed through the python mirror)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from arm_spmv_b200 import host as H

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
path = "/tmp/lap5_%d.mtx" % n
A = H.lap5_coo(n)
ri, ci, va = A.row_ind.cpu().numpy(), A.col_ind.cpu().numpy(), A.values.cpu().numpy()
t0 = time.perf_counter()
with open(path, "w") as f:
    f.write(f"%%MatrixMarket matrix coordinate real general\n{n * n} {n * n} {len(va)}\n")
    np.savetxt(f, np.column_stack([ri + 1, ci + 1, va]), fmt="%d %d %.17g")
print(f"wrote {path}: {os.path.getsize(path) / 1e6:.1f} MB in {time.perf_counter() - t0:.1f} s", flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    B = H.COOMatrixRead(path)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"GPU reader (file read + split header + parse): {1e3 * (t1 - t0):.1f} ms", flush=True)
assert np.array_equal(B.row_ind.cpu().numpy(), ri) and np.array_equal(B.col_ind.cpu().numpy(), ci)
assert np.array_equal(B.values.cpu().numpy().view(np.uint64), va.view(np.uint64))
rows, cols, nz, body = H.mtx_split(path)
import ctypes as C
from arm_spmv_b200.lib import check, current_stream, load, ptr
status = C.c_int(1)
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    check(load().thsp_mtx_parse_coo(C.c_char_p(body), C.c_size_t(len(body)), nz, ptr(B.row_ind), ptr(B.col_ind), ptr(B.values), C.byref(status), current_stream()))
    print(f"thsp_mtx_parse_coo alone ({len(body) / 1e6:.1f} MB from pageable host memory): {1e3 * (time.perf_counter() - t0):.2f} ms, status {status.value}", flush=True)
print("ok")
