"""Fixed cost of a launch of the persistent stream kernel between other kernels: power iteration on one GPU on small
grids, phase trace (THSP_CARVEOUT=0/1 compares the shared-memory carve-out hint on the vector kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import power

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
ops = power.CudaOps(dev)
for n in (40, 64, 128):
    A = power.PartitionedCSR.stencil27(n, 0, 1, ops)
    it = power.PowerIteration(A, ops)
    for _ in range(20):
        it.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        it.step()
    b.record(); torch.cuda.synchronize()
    it.trace_on()
    for _ in range(50):
        it.step()
    rep = it.trace_report()
    print(f"n={n:4d} rows={n**3:9d} step {a.elapsed_time(b) / 200 * 1e3:8.1f} us  | " + "  ".join(f"{k}: {v * 1e3:.1f} us" for k, v in rep.items()), flush=True)
