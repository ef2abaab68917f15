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


def ncu_traffic(n):
    """DRAM bytes of one launch of the timed kernel from the committed ncu capture (profiles/traffic.json, written by
    scripts/record_traffic.py) - quoted only while the kernel's sources are the ones that were profiled: a capture of
    an older kernel is reported as null, not as a number that silently went stale."""
    import hashlib
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        h = hashlib.sha256()
        for p in t["sources"]:
            h.update(open(os.path.join(ROOT, p), "rb").read())
        if h.hexdigest() != t["sources_sha256"]:
            return None, f"profiles/{t['capture']} was taken from an older version of {', '.join(t['sources'])}: not quoted"
        if t["grid"] != n:
            return None, f"the capture is of the {t['grid']}^3 matrix"
        return int(t["dram_bytes_per_launch"]), f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch ({t['capture']}, same sources by SHA-256); not measurable live"
    except (OSError, KeyError, ValueError):
        return None, "no capture recorded (profiles/traffic.json)"

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
def ref_checker():
    """(checker, kind): the unmodified reference (oracle/_ref/libref.so) with all host threads - its CSR / ELL / DIA
    loops give every row to one thread, so the bits do not depend on the thread count - else the C restatement."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    try:
        R = pyoracle.Ref()
        R.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
        return R, "reference"
    except (FileNotFoundError, OSError):
        return pyoracle.Oracle(), "port"


def cpu_reference_arm(n, reps, warm=1, iterated=False, want_y=None):
    """The reference's CPU CSR SpMV (main.cpp:54-61 protocol) on an n^3 27-point stencil - or, with
    `iterated`, the power-iteration step composed from the reference's own calls (Fill, CSRMatrixMatVector,
    vec_dot, vec_axpby: SURVEY.md 3.5), which is what the N > 1 arm measures.
    Returns (gflops, seconds_per_call, kind, cores, sample, tuned) - `tuned` describes the same measurement with
    the reference rebuilt at -O3 -march=x86-64-v3 (None when that library or the CPU features are missing).
    `want_y` (a dict) receives the parity material of the same matrix: "y" = 0 + A x computed by the same checker,
    "x" and the matrix "arrays", so that the caller can show the GPU multiplied the same matrix."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    O = pyoracle.Oracle()
    tuned = None
    rp, ci, va = O.gen_stencil27_csr(n)
    N = n ** 3
    x = O.gen_vector(N, 11)
    if want_y is not None:
        want_y["x"] = x
        want_y["arrays"] = (rp, ci, va)
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
        if want_y is not None:
            want_y["y"] = R.csr_spmv(N, N, rp, ci, va, x, np.zeros(N))
            want_y["kind"] = "reference"
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
        if want_y is not None:
            want_y["y"] = O.csr_spmv(N, N, rp, ci, va, x, np.zeros(N))
            want_y["kind"] = "port"
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


def n1_config(n, nnz=None):
    """`config` of the N = 1 line, the same dict in both arms (ours and --impl reference): the workload, not how an arm runs it."""
    N = n ** 3
    nnz = (3 * n - 2) ** 3 if nnz is None else int(nnz)
    nbytes = nnz * 12 + (N + 1) * 4 + 3 * N * 8
    return {"workload": f"fp64 CSR SpMV (y += A x), 27-point stencil {n}^3 (BASELINE configs[1])", "rows": N, "nnz": nnz,
            "cache": f"inputs ({nbytes / 1e9:.1f} GB) larger than the GPU's L2 (126 MB) and every host cache"}


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
        "config": {"workload": workload, "grid": n, "rows": N} if multi else n1_config(n),
        "arm_detail": {"what": workload, "grid": n},
        "cpu_baseline": {"value": round(gf, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(gf, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if tuned:
        line["cpu_baseline"]["rebuilt_o3"] = tuned
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_single(args):
    import numpy as np
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
    traffic, traffic_note = ncu_traffic(n) if kname == "stream" else (None, "the capture is of the stream kernel")

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
    # ---- parity material (compared with the reference's own CSRMatrixMatVector further down, where the CPU arm runs)
    parity = {}
    do_parity = not args.no_cpu and args.cpu_grid == n
    if do_parity:
        y_dev_np = yd.cpu().numpy()          # 0 + A x by the kernel that was timed
        y_host_np = yh.numpy().copy()        # the same through the host-buffer call
        gpu_arrays = (A.row_ptr.cpu().numpy(), A.col_ind.cpu().numpy(), A.values.cpu().numpy(), x.values.cpu().numpy())

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
        if do_parity:
            chk, _ = ref_checker()
            ye.Fill(0.0)
            H.ELLMatrixMatVector(E, x, ye)
            torch.cuda.synchronize()
            want = chk.ell_spmv(N, N, 27, E.col_ind.cpu().numpy(), E.values.cpu().numpy(), gpu_arrays[3], np.zeros(N))
            parity["ell_bits"] = bool(ye.values.cpu().numpy().tobytes() == want.tobytes())
            del want
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
        if do_parity:
            yd2.Fill(0.0)
            H.DIAMatrixMatVector(Dm, x, yd2)
            torch.cuda.synchronize()
            want = chk.dia_spmv(N, N, Dm.offsets.cpu().numpy(), Dm.values.cpu().numpy(), gpu_arrays[3], np.zeros(N))
            parity["dia_bits"] = bool(yd2.values.cpu().numpy().tobytes() == want.tobytes())
            del want
        del Dm, yd2, Bc

    # ---- the iterated loop (power iteration) on one GPU: the denominator of the multi-GPU runs
    if args.iterated_grid:
        from arm_spmv_b200 import power
        A.free_plan()
        del A, x, y, yd, xh, yh
        torch.cuda.empty_cache()
        g = args.iterated_grid
        # same warm-up and step count as the N > 1 runs use: their norm and fingerprints of y must equal these bit for bit
        Ap, res = power.measure(g, 0, 1, torch.device("cuda", 0), args.steps, args.warmup, ["deferred", "allgather"], True, ClockSampler,
                                with_e2e=False)
        nnz_g = (3 * g - 2) ** 3

        def loop_line(r, what):
            return {"workload": f"power iteration ({what}), 27-point stencil {g}^3 ({nnz_g} nnz, {len(Ap.blocks)} row blocks of < 2^31 entries), 1 GPU",
                    "ms_per_step": round(r["ms_per_step"], 4), "gflops": round(2.0 * nnz_g / (r["ms_per_step"] * 1e-3) / 1e9, 2),
                    "achieved_gbs": round(r["bytes_per_step_rank"] / (r["ms_per_step"] * 1e-3) / 1e9, 1),
                    "frac": round(r["bytes_per_step_rank"] / (r["ms_per_step"] * 1e-3) / 1e9 / peak, 4), "norm": r["norm"],
                    "steps_done": r["steps_done"], "y_hash": r["y_hash"]}

        eager = res["allgather"]
        extra["iterated_eager"] = loop_line(eager, "SpMV, sum of squares, normalising pass")
        lazy = res.get("deferred", {})
        if "ms_per_step" in lazy:   # the normalisation folded into the next SpMV: same bits, one pass over the vector less
            extra["iterated"] = loop_line(lazy, "normalisation deferred into the next SpMV")
            extra["iterated"]["equals_eager_bits"] = bool(lazy["norm"] == eager["norm"] and lazy["y_hash"] == eager["y_hash"])
        else:
            extra["iterated"] = extra["iterated_eager"]
            extra["iterated_deferred"] = lazy
        del Ap
        torch.cuda.empty_cache()

    cpu = None
    if not args.no_cpu:
        ref = {} if do_parity else None
        gf, dt, kind, cores, sample, tuned = cpu_reference_arm(args.cpu_grid, args.cpu_reps, want_y=ref)
        cpu = {"value": round(gf, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        if tuned:
            cpu["rebuilt_o3"] = tuned
        if do_parity:
            # the GPU multiplied the matrix and the vector the reference multiplied (device generators vs the oracle's CPU
            # twins, byte for byte), and y = 0 + A x has the reference's bits - through the device call and the host-buffer call
            parity["inputs_bits"] = bool(all(g.tobytes() == h.tobytes() for g, h in zip(gpu_arrays, ref["arrays"] + (ref["x"],))))
            parity["csr_bits"] = bool(y_dev_np.tobytes() == ref["y"].tobytes())
            parity["csr_host_buffer_bits"] = bool(y_host_np.tobytes() == ref["y"].tobytes())
            parity["checker"] = ("the reference's own CSRMatrixMatVector / ELLMatrixMatVector / DIAMatrixMatVector (oracle/_ref/libref.so)"
                                 if ref["kind"] == "reference" else "oracle/oracle.c (libref.so absent)")
            parity["what"] = f"y = 0 + A x on the {n}^3 stencil, x = gen_vector(seed 11), all {N} rows compared byte for byte"
            del gpu_arrays, ref

    line = {
        "metric": METRIC, "value": round(gflops, 2), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": round(ms, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": n1_config(n, nnz),
        "arm_detail": {"matrix": "generated on device", "kernel": kname, "lanes": lanes,
                       "step_ms_min": round(min(per), 5), "step_ms_max": round(max(per), 5)},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "algorithmic_bytes_per_launch": bytes_alg, "frac_of_8TBs_spec": round(achieved / 8000.0, 4)},
        "e2e": {"value": round(e2e_gflops, 2), "unit": UNIT, "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 8,
                "ms_per_step": round(e2e_ms, 4), "call": "thsp_csr_plan_spmv_host_f64 (pinned x in, y out)"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if parity:
        line["parity"] = parity
    line.update(extra)
    print(json.dumps(line), flush=True)
    bad = [k for k, v in parity.items() if k.endswith("_bits") and not v]
    if bad:
        print(f"bench.py: parity FAILED against the reference: {bad}", file=sys.stderr, flush=True)
        sys.exit(1)


def preflight_multi(rank, world, device):
    """Before the timed N-GPU runs: the two partitioned paths that the power iteration does not exercise, on a small matrix,
    against the CPU checker (oracle/ - checker use only).  (1) power.ColumnPartitionedCSC: CSCMatrixMatVectorNuma's column
    blocks (src/mat_vec.cpp:299-366) + the reduce-scatter of the private y's the reference leaves out; (2) the C++
    *MatVectorNuma entry points with G = N GPUs in one process (bin/api_check on a committed fixture, rank 0 only; the
    other ranks wait).  Returns a dict for the JSON line; a failure is reported there and makes bench.py exit non-zero."""
    import subprocess
    import tempfile

    import numpy as np
    import torch
    import torch.distributed as dist

    from arm_spmv_b200 import host as H, power
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    O = pyoracle.Oracle()
    out = {}
    # ---- (1) column-partitioned CSC
    n, nnz = 6000, 90000
    ri, ci, va = O.gen_uniform_coo(n, n, nnz, 47)
    cp, ro, vo = O.coo2csc(n, n, ri, ci, va)
    xh = O.gen_vector(n, 5) - 0.5
    want = O.coo_spmv(n, n, ri, ci, va, xh, np.zeros(n))
    scale = np.bincount(ri, weights=np.abs(va * xh[ci]), minlength=n)
    ops = power.CudaOps(device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    P = power.ColumnPartitionedCSC(n, n, t(cp), t(ro), t(vo), rank, world, ops)
    y = torch.zeros(P.nrow_local, dtype=torch.float64, device=device)
    P.spmv(t(xh[P.c0:P.c0 + P.ncol_local]), y)
    got = y.cpu().numpy()
    sl = slice(P.r0, P.r0 + P.nrow_local)
    den = np.maximum(scale[sl], 1e-300)
    err = torch.tensor([float(np.max(np.abs(got - want[sl]) / den)) if len(got) else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    out["csc_column_blocks_reduce_scatter"] = {"ok": bool(err.item() <= 1e-12), "max_row_error": float(err.item()), "tolerance": 1e-12,
                                               "matrix": f"uniform {n}x{n}, {nnz} entries", "gpus": world}
    # ---- (1b) the same at a size where time means something: 4M x 4M, 64 M entries, column blocks + reduce-scatter over
    # NVLink against the whole matrix on one GPU (CSCMatrixMatVector) - the hardware number SURVEY.md 8(f)-3 asks for
    try:
        nb, nzb = 1 << 22, 1 << 26
        Ab = H.uniform_coo(nb, nb, nzb, 43, device=device)
        Cb = H.CSCMatrix(Ab)
        del Ab
        xb = H.gen_vector(nb, 3, device=device)
        Pb = power.ColumnPartitionedCSC(nb, nb, Cb.col_ptr, Cb.row_ind, Cb.values, rank, world, ops)
        yb = torch.zeros(Pb.nrow_local, dtype=torch.float64, device=device)
        xs = xb.values[Pb.c0:Pb.c0 + Pb.ncol_local].clone()

        def timed(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        ms_part = timed(lambda: Pb.spmv(xs, yb))
        yw = H.Vector(nb, device=device)
        yw.Fill(0.0)
        ms_one = timed(lambda: H.CSCMatrixMatVector(Cb, xb, yw))   # every rank runs the whole matrix alone: max = one GPU's time
        out["csc_column_blocks_timing"] = {"matrix": f"uniform {nb}x{nb}, {nzb} entries", "gpus": world, "ms_per_spmv": round(ms_part, 4),
                                           "ms_one_gpu_whole_matrix": round(ms_one, 4), "speedup": round(ms_one / ms_part, 3),
                                           "note": "y = A x with the column blocks on the GPUs and one NCCL reduce-scatter of the partial y's per product"}
        del Cb, Pb, xb, yb, yw, xs
        torch.cuda.empty_cache()
    except Exception as e:
        out["csc_column_blocks_timing"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    # ---- (2) the C++ *MatVectorNuma calls on G = world GPUs (one process drives all of them)
    res = {"ok": None}
    if rank == 0:
        exe = os.path.join(ROOT, "bin", "api_check")
        gold = os.path.join(ROOT, "tests", "golden", "api_rand90")
        if os.path.exists(exe):
            with tempfile.TemporaryDirectory() as tmp:
                r = subprocess.run([exe, os.path.join(gold, "matrix.mtx"), tmp, str(world)], capture_output=True, text=True, timeout=300)
                ld = lambda d, f: np.fromfile(os.path.join(d, f), dtype=np.float64)
                ok = r.returncode == 0
                detail = {}
                if ok:
                    one = ld(gold, "y_csr_copy.f64")          # y of the unmodified reference build (fixture)
                    acc = np.zeros_like(one)
                    for _ in range(50):                        # NTESTS accumulations (src/mat_vec.cpp:270)
                        acc = acc + one
                    detail["csr_bits"] = bool(ld(tmp, "y_csr_numa.f64").tobytes() == acc.tobytes())
                    for f in ("ell", "coo", "csc"):
                        detail[f + "_close"] = bool(np.max(np.abs(ld(tmp, f"y_{f}_numa.f64") - acc)) <= 1e-11 * max(1.0, float(np.max(np.abs(acc)))))
                    ok = all(detail.values())
                res = {"ok": bool(ok), "gpus": world, "fixture": "tests/golden/api_rand90 (y from the reference build)", **detail}
                if r.returncode != 0:
                    res["stderr"] = r.stderr[-300:]
        else:
            res = {"ok": None, "skipped": "bin/api_check not built"}
    dist.barrier()
    out["cpp_matvec_numa"] = res
    return out


def run_multi(args):
    from arm_spmv_b200 import power
    pre = None
    if not args.no_preflight and env_int("WORLD_SIZE", 1) > 1:
        import torch
        import torch.distributed as dist
        local = env_int("LOCAL_RANK", 0)
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
        try:
            pre = preflight_multi(env_int("RANK", 0), env_int("WORLD_SIZE", 1), device)
        except Exception as e:
            pre = {"error": f"{type(e).__name__}: {e}"[:300]}
    rc = power.bench_main(args, METRIC, UNIT, csr_bytes, peak_hbm, ClockSampler, extra={"preflight": pre} if pre else None)
    bad = pre is not None and ("error" in pre or any(isinstance(v, dict) and v.get("ok") is False for v in pre.values()))
    if bad or rc:
        sys.exit(1)


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
    ap.add_argument("--exchange", default="xchgd,xchg,allgather,cepush,halo",
                    help="x refresh modes to time at N>1 (also: cepush, push, fused); the first that works is `value`")
    ap.add_argument("--no-overlap", action="store_true", help="do not overlap interior rows with the x refresh")
    ap.add_argument("--no-preflight", action="store_true", help="N>1: skip the small partitioned-path checks before the timed runs")
    ap.add_argument("--no-one-gpu", action="store_true", help="N>1: do not time the same workload on rank 0's GPU alone first")
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
