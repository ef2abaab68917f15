"""Host<->device copy bandwidth of this box (pinned memory), one direction at a time and both at once:
the floor under the host-buffer SpMV call (bench.py e2e)."""
import torch

n = 1 << 24  # 16.8 M doubles = 134 MB, the x / y of the 256^3 stencil
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


gb = n * 8 / 1e9
t = timed(lambda: d_in.copy_(h_in, non_blocking=True)); print(f"H2D  {t:.3f} ms  {gb / t * 1e3:.1f} GB/s")
t = timed(lambda: h_out.copy_(d_out, non_blocking=True)); print(f"D2H  {t:.3f} ms  {gb / t * 1e3:.1f} GB/s")
t = timed(both); print(f"both {t:.3f} ms  {gb / t * 1e3:.1f} GB/s each way")
