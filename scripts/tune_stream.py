"""Sweep the STREAM kernel's knobs (warps, stages, chunk, ctas) on the 256^3 stencil."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arm_spmv_b200 import host as H
from arm_spmv_b200.lib import check, current_stream, load, ptr

lib = load()
torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = n ** 3
A = H.stencil27_csr(n)
x = H.gen_vector(N, 11)
y = H.Vector(N); y.Fill(0.0)
plan = A.plan()
nbytes = A.nnz * 12 + (N + 1) * 4 + N * 8 * 3
cfgs = []
for line in sys.argv[2:]:
    cfgs.append(tuple(int(v) for v in line.split(",")))
if not cfgs:
    cfgs = [(0, 0, 0, 0), (16, 1, 896, 148), (16, 1, 1024, 148), (18, 1, 896, 148), (20, 1, 896, 148), (16, 2, 448, 148), (24, 1, 448, 148),
            (24, 2, 320, 148), (24, 1, 640, 148), (12, 1, 896, 296), (8, 1, 896, 296), (10, 1, 896, 296), (8, 2, 896, 148)]
st = current_stream()
for (w, s, c, g) in cfgs:
    check(lib.thsp_csr_plan_set_stream_config(plan, w, s, c, g))
    try:
        for _ in range(3):
            check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(y.values), 1, st))
    except Exception as e:
        print((w, s, c, g), "ERR", e); continue
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(y.values), 1, st))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"warps={w:2d} stages={s} chunk={c:4d} ctas={g:3d}  {ms:.4f} ms  {nbytes / ms / 1e6:7.1f} GB/s  {2 * A.nnz / ms / 1e6:7.1f} GFLOP/s", flush=True)
