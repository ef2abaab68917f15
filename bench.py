#!/usr/bin/env python
"""bench.py -- the reference's headline measurement on B200: SpMV GFLOP/s + achieved HBM GB/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1   workload = BASELINE.json configs[1]: fp64 CSR SpMV, 27-point stencil on 256^3
        (16.8 M rows, 449 M entries, 5.86 GB streamed per SpMV).  One step = one y += A x
        through the C ABI (thsp_csr_plan_spmv_f64) on torch's current stream.
N > 1   workload = configs[4]: row-partitioned power iteration on 512^3 (3.61 G entries), one
        process per GPU (torchrun), x replicated; one step = SpMV + sum of squares over all ranks +
        normalise + refresh of the replicas of x.  Every x-refresh mode of arm-spmv_b200/power.py
        is timed (all give bit-identical y); `value` is the first of --exchange that works - by
        default the flag-based NVLink exchange (csrc/exchange.cu), with the NCCL all-gather of the
        whole vector and the others beside it under "x_refresh_modes".  Fixed total problem ->
        "scaling": "strong".

`value` is GFLOP/s (2 nnz per SpMV) with everything resident in HBM; `e2e` is the same metric
through the host-buffer C-ABI call (pinned x in, y out, copies inside the timed region);
`roofline` is algorithmic bytes / CUDA-event time of the SpMV kernel against the measured copy
bandwidth in MEASURED_PEAKS.json; `cpu_baseline` times the reference's own CPU code
(oracle/_ref/libref.so, built from /root/reference; the C port in oracle/ if that is absent) on
this box's host cores.  --impl reference prints that CPU arm as its own JSON line (N = 1: the SpMV; N > 1: the
power-iteration step composed from the reference's own calls, on 256^3 - 512^3 does not fit its int32 structs).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "csr_spmv_gflops"
UNIT = "GFLOP/s"
# DRAM bytes of one csr_stream_kernel<double> launch on the 256^3 stencil, from the committed ncu capture
# (5.751718 GB read + 115.587 MB written) - 1.0007x the algorithmic 5,863,223,204 bytes.
NCU_TRAFFIC_BYTES = 5751718000 + 115587072
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def csr_bytes(nrow, ncol, nnz, accumulate=True, vbytes=8):
    """nnz (V+4) + (nrow+1) 4 + ncol V (x once) + nrow V (y write) [+ nrow V (y read)]."""
    return nnz * (vbytes + 4) + (nrow + 1) * 4 + ncol * vbytes + nrow * vbytes * (2 if accumulate else 1)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled every 100 ms while the timed region runs (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
def cpu_reference_arm(n, reps, warm=1, iterated=False):
    """The reference's CPU CSR SpMV (main.cpp:54-61 protocol) on an n^3 27-point stencil - or, with
    `iterated`, the power-iteration step composed from the reference's own calls (Fill, CSRMatrixMatVector,
    vec_dot, vec_axpby: SURVEY.md 3.5), which is what the N > 1 arm measures.
    Returns (gflops, seconds_per_call, kind, cores, sample, tuned) - `tuned` describes the same measurement with
    the reference rebuilt at -O3 -march=x86-64-v3 (None when that library or the CPU features are missing)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    O = pyoracle.Oracle()
    tuned = None
    rp, ci, va = O.gen_stencil27_csr(n)
    N = n ** 3
    x = O.gen_vector(N, 11)
    what = "power-iteration steps (y=0; y+=Ax; sqrt(dot); axpby)" if iterated else "back-to-back y+=Ax"
    try:
        R = pyoracle.Ref()

        def run(k):
            return R.time_power_iteration(N, rp, ci, va, x, k)[0] if iterated else R.time_csr_spmv(N, N, rp, ci, va, x, k)

        # Give the reference every host thread that helps: containers often expose more CPUs
        # than their quota serves, and the reference's static OpenMP loop then slows down with
        # threads.  Try 1, 2, 4, ... nproc on a short run and keep the fastest.
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        cand, t = [], 1
        while t < ncpu:
            cand.append(t)
            t *= 2
        cand.append(ncpu)
        best = None
        for t in cand:
            R.set_threads(t)
            run(1)
            d = run(2)
            if best is None or d < best[0]:
                best = (d, t)
        cores = best[1]
        R.set_threads(cores)
        for _ in range(warm):
            run(1)
        dt = run(reps)
        kind = "reference"
        try:   # the same sources at -O3 with AVX2/FMA, so the stock -O2 build is not a handicapped baseline
            if pyoracle.RefO3.runnable():
                R = pyoracle.RefO3()
                R.set_threads(cores)
                run(1)
                dt3 = run(max(1, reps // 2))
                tuned = {"value": round(2.0 * int(rp[-1]) / dt3 / 1e9, 4), "unit": UNIT, "cores": cores, "flags": pyoracle.RefO3.FLAGS}
        except (FileNotFoundError, OSError, AttributeError):
            tuned = None
    except (FileNotFoundError, OSError):
        y = np.zeros(N)
        t0 = time.perf_counter()
        k = max(1, reps // 4)
        for _ in range(k):
            y = O.csr_spmv(N, N, rp, ci, va, x, np.zeros(N) if iterated else y)
            if iterated:
                x = O.axpby(1.0 / np.sqrt(O.dot(y, y)), y, 0.0, y)
        dt = (time.perf_counter() - t0) / k
        kind, cores = "port", 1
    nnz = int(rp[-1])
    sample = f"27-pt stencil {n}^3 ({N} rows, {nnz} nnz), {reps} {what}, mean"
    return 2.0 * nnz / dt / 1e9, dt, kind, cores, sample, tuned


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    n = args.grid
    multi = args.gpus > 1 or env_int("WORLD_SIZE", 1) > 1
    gf, dt, kind, cores, sample, tuned = cpu_reference_arm(n, max(1, args.steps), warm=max(1, min(args.warmup, 2)), iterated=multi)
    N = n ** 3
    if multi:
        workload = (f"fp64 CSR power iteration composed from the reference's calls, 27-point stencil {n}^3, CPU reference ({kind}); "
                    "the 512^3 matrix of the GPU arm (3.6e9 entries) does not fit the reference's int32 structs, so the sample is 256^3")
    else:
        workload = f"fp64 CSR SpMV, 27-point stencil {n}^3, CPU reference ({kind})"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gf, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 4), "higher_is_better": True,
        "scaling": "strong" if multi else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "grid": n, "rows": N},
        "cpu_baseline": {"value": round(gf, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(gf, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if tuned:
        line["cpu_baseline"]["rebuilt_o3"] = tuned
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_single(args):
    import torch

    import arm_spmv_b200 as pkg
    from arm_spmv_b200 import host as H
    from arm_spmv_b200.lib import check, current_stream, launch_count, load, ptr

    torch.cuda.set_device(0)
    lib = load()
    n = args.grid
    N = n ** 3
    A = H.stencil27_csr(n)
    nnz = A.nnz
    x = H.gen_vector(N, 11)
    y = H.Vector(N)
    y.Fill(0.0)
    plan = A.plan()
    if args.kernel:
        kid = {"scalar": 1, "vector": 2, "stream": 3, "merge": 4}[args.kernel]
        check(lib.thsp_csr_plan_set_kernel(plan, kid, args.lanes))
    if args.stream_cfg:
        w, s, c, g = (int(v) for v in args.stream_cfg.split(","))
        check(lib.thsp_csr_plan_set_stream_config(plan, w, s, c, g))
    kname, lanes = A.plan_kernel()
    stream = current_stream()

    def step():
        check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(y.values), 1, stream))

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = launch_count()
    with ClockSampler(0) as clk:
        ev[0].record()
        for i in range(args.steps):
            step()
            ev[i + 1].record()
        torch.cuda.synchronize()
    launches = launch_count() - l0
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    ms = total_ms / args.steps
    gflops = 2.0 * nnz / (ms * 1e-3) / 1e9
    bytes_alg = csr_bytes(N, N, nnz, True)
    achieved = bytes_alg / (ms * 1e-3) / 1e9
    peak, peak_kind = peak_hbm()

    # ---- end to end: host x in, host y out, through the C ABI (copies inside the timed region)
    xh = torch.empty(N, dtype=torch.float64).pin_memory()
    yh = torch.empty(N, dtype=torch.float64).pin_memory()
    xh.copy_(x.values.cpu())
    xd = torch.empty(N, dtype=torch.float64, device="cuda")
    yd = torch.empty(N, dtype=torch.float64, device="cuda")
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        check(lib.thsp_csr_plan_spmv_host_f64(plan, C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()), ptr(xd), ptr(yd), 0, stream))

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns when y_host is complete
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_gflops = 2.0 * nnz / (e2e_ms * 1e-3) / 1e9
    # the e2e result is the oracle-checkable one: y_host == y of an accumulate-free device run
    check(lib.thsp_csr_plan_spmv_f64(plan, ptr(x.values), ptr(yd), 0, stream))
    torch.cuda.synchronize()
    assert torch.equal(yd.cpu(), yh), "host-buffer path and device path disagree"

    # ---- ELL on the same matrix (configs[1] is "ELL vs CSR")
    extra = {}
    if not args.no_ell:
        del xd
        E = H.stencil27_ell(n)
        ye = H.Vector(N)
        ye.Fill(0.0)
        for _ in range(3):
            H.ELLMatrixMatVector(E, x, ye)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            H.ELLMatrixMatVector(E, x, ye)
        b.record()
        torch.cuda.synchronize()
        ell_ms = a.elapsed_time(b) / args.steps
        ell_bytes = N * 27 * 12 + N * 8 + 2 * N * 8
        extra["ell"] = {"ms_per_step": round(ell_ms, 4), "gflops": round(2.0 * nnz / (ell_ms * 1e-3) / 1e9, 2),
                        "achieved_gbs": round(ell_bytes / (ell_ms * 1e-3) / 1e9, 1),
                        "frac": round(ell_bytes / (ell_ms * 1e-3) / 1e9 / peak, 4)}
        del E, ye
        # DIA: no index stream at all (8 B per entry) - the format with the highest roofline on a stencil
        Bc = H.CSRMatrix(nrow=N, ncol=N, row_ptr=A.row_ptr, col_ind=A.col_ind, values=A.values)
        Dm = H.DIAMatrix(Bc)
        yd2 = H.Vector(N)
        yd2.Fill(0.0)
        for _ in range(3):
            H.DIAMatrixMatVector(Dm, x, yd2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            H.DIAMatrixMatVector(Dm, x, yd2)
        b.record()
        torch.cuda.synchronize()
        dia_ms = a.elapsed_time(b) / args.steps
        dia_bytes = N * Dm.ndiags * 8 + Dm.ndiags * 4 + N * 8 + 2 * N * 8
        extra["dia"] = {"ms_per_step": round(dia_ms, 4), "gflops": round(2.0 * nnz / (dia_ms * 1e-3) / 1e9, 2), "ndiags": Dm.ndiags,
                        "achieved_gbs": round(dia_bytes / (dia_ms * 1e-3) / 1e9, 1), "frac": round(dia_bytes / (dia_ms * 1e-3) / 1e9 / peak, 4)}
        del Dm, yd2, Bc

    # ---- the iterated loop (power iteration) on one GPU: the denominator of the multi-GPU runs
    if args.iterated_grid:
        from arm_spmv_b200 import power
        A.free_plan()
        del A, x, y, yd, xh, yh
        torch.cuda.empty_cache()
        g = args.iterated_grid
        Ap, res = power.measure(g, 0, 1, torch.device("cuda", 0), min(args.steps, 20), 3, ["allgather"], True, ClockSampler, with_e2e=False)
        r = res["allgather"]
        nnz_g = (3 * g - 2) ** 3
        extra["iterated"] = {"workload": f"power iteration, 27-point stencil {g}^3 ({nnz_g} nnz, {len(Ap.blocks)} row blocks of < 2^31 entries), 1 GPU",
                             "ms_per_step": round(r["ms_per_step"], 4), "gflops": round(2.0 * nnz_g / (r["ms_per_step"] * 1e-3) / 1e9, 2),
                             "achieved_gbs": round(r["bytes_per_step_rank"] / (r["ms_per_step"] * 1e-3) / 1e9, 1),
                             "frac": round(r["bytes_per_step_rank"] / (r["ms_per_step"] * 1e-3) / 1e9 / peak, 4), "norm": r["norm"]}
        del Ap
        torch.cuda.empty_cache()

    cpu = None
    if not args.no_cpu:
        gf, dt, kind, cores, sample, tuned = cpu_reference_arm(args.cpu_grid, args.cpu_reps)
        cpu = {"value": round(gf, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        if tuned:
            cpu["rebuilt_o3"] = tuned

    line = {
        "metric": METRIC, "value": round(gflops, 2), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": round(ms, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"fp64 CSR SpMV (y += A x), 27-point stencil {n}^3 generated on device (BASELINE configs[1])",
                   "rows": N, "nnz": nnz, "kernel": kname, "lanes": lanes, "cache": "inputs (5.9 GB) larger than L2 (126 MB)",
                   "step_ms_min": round(min(per), 5), "step_ms_max": round(max(per), 5)},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": NCU_TRAFFIC_BYTES if (n == 256 and kname == "stream") else None,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one csr_stream_kernel launch "
                                       "(profiles/r01_ncu_stream_final.txt); not measurable live", "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "algorithmic_bytes_per_launch": bytes_alg, "frac_of_8TBs_spec": round(achieved / 8000.0, 4)},
        "e2e": {"value": round(e2e_gflops, 2), "unit": UNIT, "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 8,
                "ms_per_step": round(e2e_ms, 4), "call": "thsp_csr_plan_spmv_host_f64 (pinned x in, y out)"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if cpu:
        line["cpu_baseline"] = cpu
    line.update(extra)
    print(json.dumps(line), flush=True)


def run_multi(args):
    from arm_spmv_b200 import power
    power.bench_main(args, METRIC, UNIT, csr_bytes, peak_hbm, ClockSampler)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=None, help="stencil grid edge (default 256 at N=1, 512 at N>1)")
    ap.add_argument("--kernel", default=None, choices=[None, "scalar", "vector", "stream", "merge"])
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--stream-cfg", default=None, help="warps,stages,chunk,ctas for the stream kernel (0 = keep)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ell", action="store_true")
    ap.add_argument("--cpu-grid", type=int, default=256)
    ap.add_argument("--cpu-reps", type=int, default=20)
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="N>1: SMs the interior SpMV leaves free for the concurrent NCCL all-gather (-1 = 16*log2(N): 16/32/48)")
    ap.add_argument("--exchange", default="xchg,allgather,halo",
                    help="x refresh modes to time at N>1 (also: cepush, push, fused); the first that works is `value`")
    ap.add_argument("--no-overlap", action="store_true", help="do not overlap interior rows with the x refresh")
    ap.add_argument("--iterated-grid", type=int, default=512, help="N=1: also time the power-iteration loop on this grid (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.grid is None:
            args.grid = 256
        return run_reference(args)
    if args.gpus > 1 or env_int("WORLD_SIZE", 1) > 1:
        if args.grid is None:
            args.grid = 512
        return run_multi(args)
    if args.grid is None:
        args.grid = 256
    run_single(args)


if __name__ == "__main__":
    main()
