"""torchrun --nproc-per-node N scripts/csc_dist_check.py: column-partitioned CSC SpMV across N GPUs (power.ColumnPartitionedCSC,
reduce-scatter of the partial y's) against the CSR SpMV of the whole matrix on every rank; prints time per SpMV."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from arm_spmv_b200 import host as H, power

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 22
A = H.uniform_coo(n, n, 1 << 26, 43)
Cc = H.CSCMatrix(A)
B = H.CSRMatrix(A)
x = H.gen_vector(n, 3)
ref = H.Vector(n); ref.Fill(0.0)
H.CSRMatrixMatVector(B, x, ref)
scale = torch.zeros(n, dtype=torch.float64, device=dev)
Babs = H.CSRMatrix(nrow=n, ncol=n, row_ptr=B.row_ptr, col_ind=B.col_ind, values=B.values.abs())
H.csr_spmv_kernel(1, 1, Babs, x.values.abs(), scale, False)
ops = power.CudaOps(dev)
P = power.ColumnPartitionedCSC(n, n, Cc.col_ptr, Cc.row_ind, Cc.values, rank, world, ops)
y = torch.zeros(P.nrow_local, dtype=torch.float64, device=dev)
xs = x.values[P.c0:P.c0 + P.ncol_local].clone()
for _ in range(3):
    P.spmv(xs, y)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    P.spmv(xs, y)
b.record(); torch.cuda.synchronize()
err = ((y - ref.values[P.r0:P.r0 + P.nrow_local]).abs() / scale[P.r0:P.r0 + P.nrow_local].clamp_min(1e-300)).max().item()
print(f"rank {rank}/{world}: column-partitioned CSC SpMV {a.elapsed_time(b) / 10:.3f} ms, max per-row error {err:.2e}", flush=True)
assert err <= 1e-12
dist.barrier(); dist.destroy_process_group()
