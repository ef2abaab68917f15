"""What a chunk costs on the PCIe pipeline of the host-buffer SpMV (bench.py e2e): 134 MB up and 134 MB down in K
pieces - one direction alone, both directions independent, y piece c waiting for x piece c, and with a small kernel
between them - submitted eagerly and replayed as a CUDA graph."""
import time
import torch

n = 1 << 24
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s0, s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def pipeline(K, mode):
    m = n // K
    ev0 = torch.cuda.Event(); ev0.record(s0)
    s1.wait_event(ev0); s2.wait_event(ev0)
    for c in range(K):
        sl = slice(c * m, (c + 1) * m)
        if mode != "d2h":
            with torch.cuda.stream(s1):
                d_in[sl].copy_(h_in[sl], non_blocking=True)
                e1 = torch.cuda.Event(); e1.record(s1)
        if mode in ("dep", "kernel"):
            s0.wait_event(e1)
            if mode == "kernel":
                with torch.cuda.stream(s0):
                    torch.add(d_in[sl], 1.0, out=d_out[sl])
            e2 = torch.cuda.Event(); e2.record(s0)
            s2.wait_event(e2)
        if mode != "h2d":
            with torch.cuda.stream(s2):
                h_out[sl].copy_(d_out[sl], non_blocking=True)
    e3 = torch.cuda.Event(); e3.record(s1); s0.wait_event(e3)
    e4 = torch.cuda.Event(); e4.record(s2); s0.wait_event(e4)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("mode      K   eager ms   graph ms")
for mode in ("h2d", "d2h", "indep", "dep", "kernel"):
    for K in (1, 8, 32, 128):
        te = timed(lambda: pipeline(K, mode))
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s0):
            pipeline(K, mode)
        tg = timed(g.replay)
        print(f"{mode:8s} {K:4d} {te:9.3f} {tg:9.3f}", flush=True)

# Does a kernel that saturates HBM slow the copy engines?  The same independent up + down transfers (K = 8) next to
# back-to-back 1 GB device copies on a fourth stream, and next to an SM-only spin of similar length.
big_a = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
big_b = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
s3 = torch.cuda.Stream()


def with_load(reps_load, what):
    ev = torch.cuda.Event(); ev.record(s0); s3.wait_event(ev)
    with torch.cuda.stream(s3):
        for _ in range(reps_load):
            if what == "hbm":
                big_b.copy_(big_a, non_blocking=True)
            else:
                torch.cuda._sleep(600000)
    pipeline(8, "indep")
    e = torch.cuda.Event(); e.record(s3); s0.wait_event(e)


def copies_only_time(reps_load, what, reps=5):
    """Time of the copy streams alone (events on s1/s2), while the load runs beside them."""
    out = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a1.record(s0)
        with_load(reps_load, what)
        # pipeline() joined s1 and s2 into s0 before the load joined it: an event recorded now on s1/s2 is at their end
        b1.record(s1); b2 = torch.cuda.Event(enable_timing=True); b2.record(s2)
        torch.cuda.synchronize()
        out.append(max(a1.elapsed_time(b1), a1.elapsed_time(b2)))
    return sum(out[1:]) / reps


print("copies (134 MB up + 134 MB down, K = 8) alone:          %.3f ms" % copies_only_time(0, "hbm"))
print("... next to 10 x 1 GB device copies (HBM saturated):     %.3f ms" % copies_only_time(10, "hbm"))
print("... next to 4 x 1 GB device copies (1.3 ms of the run):  %.3f ms" % copies_only_time(4, "hbm"))
print("... next to a spinning kernel (SMs busy, HBM idle):      %.3f ms" % copies_only_time(10, "spin"))
