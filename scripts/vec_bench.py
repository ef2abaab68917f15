"""vec_dot / vec_axpby / Vector ops at full size against the HBM roofline (dot 2nV, axpby 3nV or 2nV).
  python scripts/vec_bench.py [n ...]   (default 16777216 and 134217728 doubles)"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200.lib import check, current_stream, load, ptr

lib = load()
torch.cuda.set_device(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def t(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n in [int(v) for v in sys.argv[1:]] or [1 << 24, 1 << 27]:
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.rand(n, dtype=torch.float64, device="cuda")
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    tiles = torch.empty((n + 31) // 32, dtype=torch.float64, device="cuda")
    s = current_stream
    rows = [
        ("vec_axpby  w = a x + b y   (3 n V)", lambda: check(lib.thsp_axpby_f64(C.c_int64(n), C.c_double(0.7), ptr(x), C.c_double(1.3), ptr(y), ptr(w), s())), 3),
        ("vec_axpby  w = a x         (2 n V)", lambda: check(lib.thsp_axpby_f64(C.c_int64(n), C.c_double(0.7), ptr(x), C.c_double(0.0), ptr(y), ptr(w), s())), 2),
        ("AddScaled  y += a x        (3 n V)", lambda: check(lib.thsp_add_scaled_f64(C.c_int64(n), C.c_double(0.7), ptr(x), ptr(y), s())), 3),
        ("vec_dot (device result)    (2 n V)", lambda: check(lib.thsp_dot_dev_f64(C.c_int64(n), ptr(x), ptr(y), ptr(out), s())), 2),
        ("dot, canonical order       (2 n V)", lambda: check(lib.thsp_dot_canonical_dev_f64(C.c_int64(n), ptr(x), ptr(y), ptr(tiles), ptr(out), s())), 2),
        ("Fill                       (1 n V)", lambda: check(lib.thsp_fill_f64(C.c_int64(n), C.c_double(1.5), ptr(w), s())), 1),
    ]
    print(f"n = {n} doubles ({n * 8 / 1e6:.0f} MB per vector)")
    for name, fn, k in rows:
        try:
            ms = t(fn)
        except AttributeError as e:
            print(f"  {name}: {e}")
            continue
        gbs = k * n * 8 / ms / 1e6
        print(f"  {name}  {ms:8.4f} ms  {gbs:7.1f} GB/s  {100 * gbs / PEAK:5.1f} % of measured HBM")
