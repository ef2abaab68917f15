"""torchrun --nproc-per-node N scripts/pcie_probe_ranks.py: host<->device copy bandwidth with ALL ranks copying at once -
what bounds the multi-GPU host-buffer call (bench.py e2e at N > 1).  Each rank moves 134 MB (its share of x / y at 512^3
on 8 GPUs) H2D, D2H and both ways together, from (a) its own cudaHostAlloc'ed buffer and (b) a window of ONE /dev/shm
segment registered with cudaHostRegister by every rank (power.SharedHostVector: where x of the shared-x call lives).
Reports per-rank and whole-box GB/s, first with one rank alone, then with all ranks together."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C

import torch
import torch.distributed as dist

from arm_spmv_b200.lib import check, load

lib = load()


def h2d_copy(dst, src, stream):   # cudaMemcpyAsync through the C ABI: torch cannot know that a registered segment is page-locked
    check(lib.thsp_memcpy_h2d(C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), C.c_size_t(dst.numel() * 8), C.c_void_p(stream.cuda_stream)))


def d2h_copy(dst, src, stream):
    check(lib.thsp_memcpy_d2h(C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), C.c_size_t(dst.numel() * 8), C.c_void_p(stream.cuda_stream)))

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 24
gb = n * 8 / 1e9
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.ones(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=8):
    fn(); torch.cuda.synchronize()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run(label, h_in, h_out, active):
    """active: which ranks copy (the others only take part in the barriers)"""
    def h2d():
        if rank in active:
            h2d_copy(d_in, h_in, s1)

    def d2h():
        if rank in active:
            d2h_copy(h_out, d_out, s2)

    def both():
        if rank in active:
            h2d_copy(d_in, h_in, s1)
            d2h_copy(h_out, d_out, s2)

    k = len(active)
    res = []
    for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both)):
        t = timed(fn)
        res.append(f"{name} {t:7.3f} ms = {gb / t * 1e3:6.1f} GB/s per rank, {k * gb / t * 1e3:7.1f} GB/s box" + (" each way" if name == "both" else ""))
    if rank == 0:
        print(f"{label:44s} " + " | ".join(res), flush=True)


own_in = torch.empty(n, dtype=torch.float64).pin_memory()
own_out = torch.empty(n, dtype=torch.float64).pin_memory()
run("own pinned buffers, rank 0 alone", own_in, own_out, {0})
if world > 1:
    run(f"own pinned buffers, all {world} ranks at once", own_in, own_out, set(range(world)))
try:
    from arm_spmv_b200 import power
    xs = power.SharedHostVector(n * world, rank, world, dev, tag="pcieprobe")
    ys = power.SharedHostVector(n * world, rank, world, dev, tag="pcieprobe_y")
    win_in = xs.tensor[rank * n:(rank + 1) * n]
    win_out = ys.tensor[rank * n:(rank + 1) * n]
    run("windows of one registered /dev/shm segment, rank 0", win_in, win_out, {0})
    if world > 1:
        run(f"windows of one registered /dev/shm segment, all {world}", win_in, win_out, set(range(world)))
    xs.close(); ys.close()
except Exception as e:   # no /dev/shm or registration refused
    if rank == 0:
        print("shared segment unavailable:", e)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
