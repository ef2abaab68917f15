"""What bounds the x gather of the R-MAT SpMV?  Time a plain gather y[i] = x[idx[i]] (torch.index_select as a
neutral gather kernel) for index streams that differ in ONE property each: footprint, skew, spatial clustering."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H

torch.cuda.set_device(0)
A = H.rmat_coo(24, 16 << 24, 42)
cols = A.col_ind.to(torch.int64)
n = cols.numel()
del A


def timed(tag, x, idx):
    out = torch.empty(idx.numel(), dtype=x.dtype, device="cuda")
    for _ in range(2):
        torch.index_select(x, 0, idx, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        torch.index_select(x, 0, idx, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{tag:58s} {ms:8.3f} ms  {idx.numel() / ms / 1e6:7.1f} G gathers/s", flush=True)


for bits in (20, 22, 23, 24, 25, 26):
    x = torch.rand(1 << bits, dtype=torch.float64, device="cuda")
    idx = torch.randint(0, 1 << bits, (n,), device="cuda", dtype=torch.int64)
    timed(f"uniform indices over {8 << bits >> 20} MB", x, idx)
    del x, idx
x = torch.rand(1 << 24, dtype=torch.float64, device="cuda")
timed("R-MAT columns (skewed, clustered at low indices)", x, cols)
scr = (cols * 0x9E3779B1) & ((1 << 24) - 1)     # odd multiplier: a bijection on 24 bits; keeps the skew, breaks the clustering
timed("R-MAT columns scrambled (same skew, no clustering)", x, scr)
srt, _ = torch.sort(cols[: n // 4])
timed("R-MAT columns sorted (quarter)", x, srt)
timed("R-MAT columns, first quarter unsorted", x, cols[: n // 4].contiguous())
