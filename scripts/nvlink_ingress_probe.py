"""torchrun --nproc-per-node N scripts/nvlink_ingress_probe.py [grid]: what full replication of x costs on this box, with
nothing else running - the ceiling under the `allgather` / `cepush` refresh modes of the power iteration (bench.py N > 1).
Every rank owns N/world doubles of a vector of grid^3 doubles (default 512^3 = 1.07 GB) that is replicated on all ranks:
  (a) NCCL all_gather_into_tensor, in place;
  (b) the copy engines: at step k rank r copies its slice into rank (r + k) mod world, the world-1 steps on world-1 streams
      (permutations: no GPU receives from two senders in one step);
  (c) the same copies one after the other on one stream (how much the parallel steps buy);
  (d) one rank alone copying its slice to one peer (a single link's rate).
Prints ms and the NVLink ingress rate per GPU = (world-1) * slice bytes / time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = g ** 3
cnt = N // world
start = rank * cnt
x = symm_mem.empty(N, dtype=torch.float64, device=dev)
hdl = symm_mem.rendezvous(x, group=dist.group.WORLD.group_name)
x.fill_(float(rank))
own = x[start:start + cnt]
peers = [hdl.get_buffer(r, (N,), torch.float64) for r in range(world)]
streams = [torch.cuda.Stream(device=dev) for _ in range(max(world - 1, 1))]
token = torch.zeros(1, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    dist.all_reduce(token)   # lines the GPUs up on the stream
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def nccl_ag():
    dist.all_gather_into_tensor(x, own)


def ce_parallel():
    cur = torch.cuda.current_stream()
    ev0 = torch.cuda.Event()
    ev0.record(cur)
    for k, st in enumerate(streams[:world - 1], start=1):
        dst = (rank + k) % world
        st.wait_event(ev0)
        with torch.cuda.stream(st):
            peers[dst][start:start + cnt].copy_(own, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        cur.wait_event(ev)
    hdl.barrier()


def ce_serial():
    for k in range(1, world):
        dst = (rank + k) % world
        peers[dst][start:start + cnt].copy_(own, non_blocking=True)
    hdl.barrier()


def one_link():
    if rank == 0:
        peers[1][start:start + cnt].copy_(own, non_blocking=True)
    hdl.barrier()


ingress = (world - 1) * cnt * 8
rows = []
for name, fn, nbytes in (("NCCL all_gather_into_tensor (in place)", nccl_ag, ingress),
                         ("copy engines, world-1 staggered permutations side by side", ce_parallel, ingress),
                         ("copy engines, the same copies one after the other", ce_serial, ingress),
                         ("one rank, one peer copy (single pair)", one_link, cnt * 8)):
    ms = timed(fn)
    rows.append((name, ms, nbytes / ms / 1e6))
if rank == 0:
    print(f"# {world} GPUs, vector of {g}^3 doubles ({N * 8 / 1e9:.2f} GB), slice {cnt * 8 / 1e6:.1f} MB; ingress per GPU per refresh {ingress / 1e9:.3f} GB")
    for name, ms, gbs in rows:
        print(f"{name:62s} {ms:8.3f} ms  {gbs:7.1f} GB/s into each GPU")
dist.barrier()
dist.destroy_process_group()
