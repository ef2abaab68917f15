"""CSR stream kernel on the uniform 8M matrix (random x gathers, configs[3]) vs warps per CTA: fewer warps = smaller
shared-memory stage = more of the 256 KB array left to L1, where the lines of outstanding gather misses live."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H
from arm_spmv_b200.lib import check, current_stream, load, ptr

lib = load()
torch.cuda.set_device(0)
which = sys.argv[1] if len(sys.argv) > 1 else "uniform"
A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43) if which == "uniform" else H.rmat_coo(24, 16 << 24, 42)
B = H.CSRMatrix(A)
del A
x = H.gen_vector(B.ncol, 3)
y = H.Vector(B.nrow); y.Fill(0.0)
plan = B.plan()
check(lib.thsp_csr_plan_set_kernel(plan, 3, 1))
for warps in (4, 6, 8, 10, 12, 16, 20, 24):
    for ctas_per_sm in (1,):
        check(lib.thsp_csr_plan_set_stream_config(plan, warps, 1, 0, 148 * ctas_per_sm))
        try:
            for _ in range(3):
                check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(y.values), 1, current_stream()))
        except Exception as e:
            print(warps, "failed", str(e)[:100]); continue
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(y.values), 1, current_stream()))
        b.record(); torch.cuda.synchronize()
        print(f"{which}: stream kernel, {warps:2d} warps per CTA: {a.elapsed_time(b) / 10:.4f} ms", flush=True)
