"""Why is the stream kernel slower on a 512^3/8 slab than on 256^3 (same rows, same entries)?  Times one rank's slab of the
512^3 stencil (rows [3 N/8, 4 N/8): both neighbours exist) and the whole 256^3 matrix, y = A x, on one GPU.
  python scripts/slab_probe.py [reps]           # under ncu: --set full -k regex:csr_stream -c 2
Env THSP_STREAM_CFG=warps,stages,chunk,ctas overrides the plan's shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H
from arm_spmv_b200.lib import check, current_stream, load, ptr

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = load()
torch.cuda.set_device(0)


def run(name, n, r0, r1):
    A = H.stencil27_csr(n, r0, r1)
    x = H.gen_vector(n ** 3, 11)
    y = torch.empty(r1 - r0, dtype=torch.float64, device="cuda")
    plan = A.plan()
    cfg = os.environ.get("THSP_STREAM_CFG")
    if cfg:
        w, s, c, g = (int(v) for v in cfg.split(","))
        check(lib.thsp_csr_plan_set_stream_config(plan, w, s, c, g))
    for _ in range(3):
        check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(y), 0, current_stream()))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(y), 0, current_stream()))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    rows = r1 - r0
    byt = A.nnz * 12 + (rows + 1) * 4 + rows * 8 + (rows + 2 * (n * n + n + 1)) * 8
    print(f"{name}: rows {rows} nnz {A.nnz} kernel {A.plan_kernel()[0]}  {ms:.4f} ms  {byt / ms / 1e6:.0f} GB/s  {2 * A.nnz / ms / 1e6:.1f} GFLOP/s", flush=True)
    A.free_plan()


N5 = 512 ** 3
which = os.environ.get("SLAB_WHICH", "both")
if which in ("both", "512"):
    run("512^3 slab 3/8", 512, 3 * N5 // 8, 4 * N5 // 8)
if which in ("both", "256"):
    run("256^3 whole", 256, 0, 256 ** 3)
