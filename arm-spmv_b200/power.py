"""Row-partitioned repeated SpMV (power iteration) across the GPUs of one box.

This is what the reference's NUMA placement (src/mat_vec.cpp:230-297, include/numa_node.h)
becomes on B200s: one process per GPU (torchrun), each owning an equal block of rows
(`rows_per_node = nrow / nparts`, last block takes the remainder, :233,245-246) with row
pointers rebased to 0 (:260-263), a full-length replica of x (:257,266) and its own slice of y.
The reference never refreshes x between its 50 repeats; the north star asks for a real iterated
loop, composed from the reference's own vector ops:

    y = A x                      CSRMatrixMatVector          (thsp_csr_plan_spmv_f64)
    s = sum y_i^2 ; all-reduce   vec_dot(y, y)               (thsp_sumsq_dev_f64 + NCCL, 8 bytes)
    x_own = y / sqrt(s)          vec_axpby(1/nrm, y, 0, ...) (thsp_scale_broadcast_f64)
    refresh the replicas of x    -- no reference counterpart --

x refresh modes
  allgather  NCCL all-gather of every slice into every replica (the north star's wording).
  fused      the scale kernel stores each normalised value straight into all replicas through
             NVLink peer pointers (torch symmetric memory): compute + "all-gather" in one kernel.
  push       the scale kernel writes the local replica only (critical path); a second kernel on the
             side stream stores that slice into the other replicas while interior rows are multiplied.
  cepush     like push, but the copy engines move the slice (one peer-to-peer cudaMemcpyAsync per
             replica): no SM is taken from the SpMV that runs meanwhile.
  halo       column-footprint analysis: a block only needs x over [min col, max col] of its rows;
             only the parts of that range owned by other ranks are pulled.
  xchg       the same footprint, but nothing is pulled and no collective is called: the kernel that
             normalises y stores the pieces other ranks read straight into their replicas and raises
             a flag; the sum of squares travels the same way (csrc/exchange.cu).  One stream, four
             kernels per step (interior rows, wait, boundary rows, everything else), no host synchronisation.
With `overlap`, rows whose columns all fall inside the own slice (interior) are multiplied while
the refresh is still in flight; the boundary blocks wait for it.

Row blocks hold fewer than 2^31 entries each, so int32 row pointers survive matrices (512^3:
3.6e9 entries) that the reference's structs cannot represent.

The arithmetic lives behind an `ops` object.  The product uses CudaOps (the C ABI).  The CPU
tests (gloo, world_size 2) inject their own ops built on the oracle - this module has no CPU path.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import time
from dataclasses import dataclass, field

import torch
import torch.distributed as dist


def partition_rows(n: int, nparts: int, part: int) -> tuple[int, int]:
    """(start, count) of block `part`: src/mat_vec.cpp:233,245-246."""
    per = n // nparts
    start = part * per
    return start, (n - start) if part == nparts - 1 else per


def stencil_row_blocks(n: int, start: int, count: int, world: int, max_rows: int) -> list[tuple[int, int, bool]]:
    """Split rows [start, start+count) of the n^3 27-point stencil into (r0, r1, boundary) pieces.
    Boundary pieces are the first / last grid plane of the slab when another rank owns the
    neighbouring plane; interior pieces are cut so that each holds < 2^31 entries.  Every piece starts a
    multiple of 32 rows after `start`, so that the pieces' 32-row tiles are tiles of the whole slice (the
    canonical sum of squares, csrc/tree_sum.cuh); boundary pieces are rounded outwards for that."""
    end = start + count
    plane = n * n
    reach = plane + n + 1            # furthest column offset of a row
    up = lambda v: (v + 31) // 32 * 32
    pieces = []
    lo, hi = 0, count                # relative to start
    if world > 1 and start > 0:
        b = min(count, up(reach))
        pieces.append((start, start + b, True))
        lo = b
    tail = None
    if world > 1 and end < n ** 3 and hi > lo:
        t0 = max(lo, (hi - reach) // 32 * 32)
        tail = (start + t0, end, True)
        hi = t0
    step = max(32, max_rows // 32 * 32)
    r = lo
    while r < hi:
        e = min(hi, r + step)
        pieces.append((start + r, start + e, False))
        r = e
    if tail:
        pieces.append(tail)
    return pieces


@dataclass
class RowBlock:
    row0: int            # first global row
    nrow: int
    nnz: int
    boundary: bool       # needs x entries owned by another rank
    payload: object      # ops-specific (device CSR + plan)
    col_min: int = 0
    col_max: int = 0


class CudaOps:
    """The product arithmetic: every call goes through libthsparse_cuda.so on torch's current stream."""

    def __init__(self, device):
        from . import host as H
        from .lib import check, current_stream, load, ptr
        self.H, self.check, self.stream, self.lib, self.ptr = H, check, current_stream, load(), ptr
        self.device = device

    def empty(self, n):
        return torch.empty(n, dtype=torch.float64, device=self.device)

    def scalar(self):
        return torch.zeros(1, dtype=torch.float64, device=self.device)

    def stencil_block(self, n, r0, r1):
        A = self.H.stencil27_csr(n, r0, r1, device=self.device)
        A.plan()
        return A, A.nnz

    def csr_block(self, nrow, ncol, row_ptr, col_ind, values):
        A = self.H.CSRMatrix(nrow=nrow, ncol=ncol, row_ptr=row_ptr, col_ind=col_ind, values=values, device=self.device)
        A.plan()
        return A, A.nnz

    def csc_block(self, nrow, ncol, col_ptr, row_ind, values):
        return self.H.CSCMatrix(nrow=nrow, ncol=ncol, col_ptr=col_ptr, row_ind=row_ind, values=values, device=self.device)

    def csc_spmv(self, payload, x, y):
        """y += A_block x through thsp_csc_spmv_f64 (x, y: torch tensors)."""
        self.check(self.lib.thsp_csc_spmv_f64(payload.nrow, payload.ncol, payload.nnz, self.ptr(payload.col_ptr), self.ptr(payload.row_ind),
                                              self.ptr(payload.values), self.ptr(x), self.ptr(y), self.stream()))

    def spmv(self, payload, x, y, tile_ss=None, xscale=None):
        """y = A_block x (overwrite: saves Fill(0) and the read of y); with tile_ss also the sum of squares of every
        32-row tile of y, written by the SpMV's epilogue (thsp_csr_plan_spmv_sumsq_f64); with xscale (device scalar)
        y = A_block (xscale * x), the factor applied to every gathered x_j inside the stream kernel."""
        if xscale is not None:
            if payload.plan_kernel()[0] == "stream":
                self.check(self.lib.thsp_csr_plan_spmv_scaled_f64(payload.plan(), self.ptr(x), self.ptr(xscale), self.ptr(y), 0,
                                                                  self.ptr(tile_ss), self.stream()))
                return
            # other kernels (short or irregular rows): the same products from a scaled copy of x
            if getattr(self, "_xs", None) is None or self._xs.numel() != x.numel():
                self._xs = self.empty(x.numel())
            self.check(self.lib.thsp_scale_by_dev_f64(C.c_int64(x.numel()), self.ptr(x), self.ptr(xscale), self.ptr(self._xs), self.stream()))
            x = self._xs
        if tile_ss is None:
            self.check(self.lib.thsp_csr_plan_spmv_f64(payload.plan(), self.ptr(x), self.ptr(y), 0, self.stream()))
        else:
            self.check(self.lib.thsp_csr_plan_spmv_sumsq_f64(payload.plan(), self.ptr(x), self.ptr(y), 0, self.ptr(tile_ss), self.stream()))

    def tree_sum(self, vals, out):
        """out[0] = the canonical (index-bit tree) sum of vals: tile partials of one rank, or the ranks' partials."""
        self.check(self.lib.thsp_tree_sum_f64(C.c_int64(vals.numel()), self.ptr(vals), self.ptr(out), self.stream()))

    def inv_sqrt(self, ss, inv):
        """inv[0] = 1 / sqrt(ss[0]), on the device"""
        self.check(self.lib.thsp_inv_sqrt_dev_f64(self.ptr(ss), self.ptr(inv), self.stream()))

    def scaled(self, x, scale):
        """a new vector scale[0] * x (what a deferred normalisation would have stored)"""
        out = self.empty(x.numel())
        self.check(self.lib.thsp_scale_by_dev_f64(C.c_int64(x.numel()), self.ptr(x), self.ptr(scale), self.ptr(out), self.stream()))
        return out

    def hash(self, v, first):
        h = C.c_uint64(0)
        self.check(self.lib.thsp_hash_f64(C.c_int64(v.numel()), self.ptr(v), C.c_uint64(first), C.byref(h), self.stream()))
        return int(h.value)

    def reserve_sms(self, payload, reserve):
        """Run this block's persistent SpMV on (SMs - reserve) CTAs so that a collective launched on
        another stream finds free SMs instead of queueing behind the persistent kernel."""
        n = C.c_int(0)
        self.check(self.lib.thsp_sm_count(C.byref(n)))
        self.check(self.lib.thsp_csr_plan_set_stream_config(payload.plan(), 0, 0, 0, max(1, n.value - reserve)))

    def sumsq(self, y, out):
        self.check(self.lib.thsp_sumsq_dev_f64(C.c_int64(y.numel()), self.ptr(y), self.ptr(out), self.stream()))

    def scale_into(self, y, sumsq, x, offset, peer_ptrs=None):
        """x_k[offset + i] = y[i] / sqrt(sumsq): into the local replica `x`, or into every replica
        in peer_ptrs (device pointers mapped over NVLink) when given."""
        dst_ptrs = peer_ptrs if peer_ptrs else [x.data_ptr()]
        arr = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
        self.check(self.lib.thsp_scale_broadcast_f64(C.c_int64(y.numel()), self.ptr(y), self.ptr(sumsq), arr, len(dst_ptrs),
                                                     C.c_int64(offset), self.stream()))

    def col_range(self, payload):
        c = payload.col_ind
        if c.numel() == 0:
            return 0, -1
        return int(c.min().item()), int(c.max().item())

    # ---- flag-based exchange over peer pointers (csrc/exchange.cu) --------------------------
    def xchg_setup(self, world, rank, group_name):
        import torch.distributed._symmetric_memory as symm_mem
        lib = self.lib
        self.xw, self.xr = world, rank
        self.ctrl = symm_mem.empty(lib.thsp_xchg_ctrl_bytes() // 8, dtype=torch.int64, device=self.device)
        self.ctrl.zero_()
        self.ctrl_h = symm_mem.rendezvous(self.ctrl, group=group_name)
        self.ctrl_ptrs = [int(self.ctrl_h.buffer_ptrs[r]) for r in range(world)]
        self.ctrl_arr = (C.c_void_p * world)(*self.ctrl_ptrs)
        self.work = torch.zeros(lib.thsp_xchg_work_bytes() // 4, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize()
        self.ctrl_h.barrier()

    def xchg_dests(self, dests, x_peer_ptrs):
        """dests = [(rank, lo, hi)]: pieces of this rank's slice that other ranks read."""
        n = len(dests)
        self.nd = n
        self.d_x = (C.c_void_p * max(n, 1))(*[x_peer_ptrs[r] for r, _, _ in dests])
        self.d_ctrl = (C.c_void_p * max(n, 1))(*[self.ctrl_ptrs[r] for r, _, _ in dests])
        self.d_lo = (C.c_int64 * max(n, 1))(*[lo for _, lo, _ in dests])
        self.d_hi = (C.c_int64 * max(n, 1))(*[hi for _, _, hi in dests])

    def xchg_dests2(self, dests, peer_ptr_sets):
        """the same pieces for each of the two vectors of the deferred loop (DeferredPowerIteration)"""
        self.xchg_dests(dests, peer_ptr_sets[0])
        n = len(dests)
        self.d_x2 = [(C.c_void_p * max(n, 1))(*[ptrs[r] for r, _, _ in dests]) for ptrs in peer_ptr_sets]

    def norm_push(self, y, tile_ss, it, xout, offset, ss, inv, which):
        """tree + publish + push of the raw pieces + flags + combine -> ss, inv (thsp_xchg_norm_push_f64); `which` = the
        vector (0 / 1) the pieces belong to on every rank; xout is this rank's copy of it (the kernel goes by pointers)"""
        self.check(self.lib.thsp_xchg_norm_push_f64(C.c_int64(y.numel()), self.ptr(y), self.ptr(tile_ss), C.c_uint64(it), self.xw, self.xr,
                                                    self.ctrl_arr, self.ptr(self.work), C.c_int64(offset), self.nd, self.d_x2[which],
                                                    self.d_ctrl, self.d_lo, self.d_hi, self.ptr(ss), self.ptr(inv), self.stream()))

    def sumsq_publish(self, y, it):
        self.check(self.lib.thsp_xchg_sumsq_publish_f64(C.c_int64(y.numel()), self.ptr(y), C.c_uint64(it), self.xw, self.xr, self.ctrl_arr,
                                                        self.ptr(self.work), self.stream()))

    def scale_push(self, y, it, x, offset, ss):
        self.check(self.lib.thsp_xchg_scale_push_f64(C.c_int64(y.numel()), self.ptr(y), C.c_uint64(it), self.xw, self.xr,
                                                     C.c_void_p(self.ctrl_ptrs[self.xr]), self.ptr(self.work), self.ptr(x), C.c_int64(offset),
                                                     self.nd, self.d_x, self.d_ctrl, self.d_lo, self.d_hi, self.ptr(ss), self.stream()))

    def norm_scale_push(self, y, tile_ss, it, x, offset, ss):
        """reduce + publish + combine + normalise + push + flags in one kernel (thsp_xchg_norm_scale_push_f64)"""
        self.check(self.lib.thsp_xchg_norm_scale_push_f64(C.c_int64(y.numel()), self.ptr(y), self.ptr(tile_ss), C.c_uint64(it), self.xw, self.xr,
                                                          self.ctrl_arr, self.ptr(self.work), self.ptr(x), C.c_int64(offset), self.nd, self.d_x,
                                                          self.d_ctrl, self.d_lo, self.d_hi, self.ptr(ss), self.stream()))

    def wait_halo(self, it, src_mask):
        if it > 0 and src_mask:
            self.check(self.lib.thsp_xchg_wait(C.c_void_p(self.ctrl_ptrs[self.xr]), C.c_uint64(it), C.c_uint(src_mask), self.stream()))

    def xchg_timed_out(self):
        f = C.c_int(0)
        self.check(self.lib.thsp_xchg_timed_out(C.c_void_p(self.ctrl_ptrs[self.xr]), C.byref(f), self.stream()))
        return bool(f.value)


def pushes_from_needs(all_needs, rank):
    """all_needs[r] = [(owner, lo, hi)] of rank r (PartitionedCSR.needed_ranges).  Returns the pieces of
    `rank`'s slice that other ranks read, one bounding range per reader: [(reader, lo, hi)]."""
    out = []
    for r, needs in enumerate(all_needs):
        mine = [(lo, hi) for owner, lo, hi in needs if owner == rank and r != rank]
        if mine:
            out.append((r, min(a for a, _ in mine), max(b for _, b in mine)))
    return out


class PartitionedCSR:
    """This rank's rows of a square matrix, as a list of RowBlocks (interior first is not required)."""

    def __init__(self, n_global: int, rank: int, world: int, start: int, count: int, blocks: list[RowBlock]):
        self.N, self.rank, self.world, self.start, self.count, self.blocks = n_global, rank, world, start, count, blocks
        self.nnz_local = sum(b.nnz for b in blocks)

    @staticmethod
    def stencil27(n: int, rank: int, world: int, ops, max_block_rows: int = 1 << 26) -> "PartitionedCSR":
        N = n ** 3
        start, count = partition_rows(N, world, rank)
        blocks = []
        for r0, r1, bnd in stencil_row_blocks(n, start, count, world, max_block_rows):
            payload, nnz = ops.stencil_block(n, r0, r1)
            reach = n * n + n + 1
            blocks.append(RowBlock(r0, r1 - r0, nnz, bnd, payload, max(0, r0 - reach), min(N - 1, r1 - 1 + reach)))
        return PartitionedCSR(N, rank, world, start, count, blocks)

    @staticmethod
    def from_csr(nrow: int, row_ptr, col_ind, values, rank: int, world: int, ops) -> "PartitionedCSR":
        """Generic square CSR given as torch tensors: slice this rank's rows and rebase row_ptr
        exactly as CSRMatrixMatVectorNuma does (src/mat_vec.cpp:245-265)."""
        start, count = partition_rows(nrow, world, rank)
        e0, e1 = int(row_ptr[start]), int(row_ptr[start + count])
        sub_rp = (row_ptr[start:start + count + 1] - e0).to(torch.int32).contiguous()
        payload, nnz = ops.csr_block(count, nrow, sub_rp, col_ind[e0:e1].contiguous(), values[e0:e1].contiguous())
        lo, hi = ops.col_range(payload)
        bnd = world > 1 and (lo < start or hi >= start + count)
        return PartitionedCSR(nrow, rank, world, start, count, [RowBlock(start, count, nnz, bnd, payload, lo, hi)])

    def spmv(self, ops, x, y_local, boundary: bool | None = None, tile_ss=None, xscale=None):
        kw = {} if xscale is None else {"xscale": xscale}
        for b in self.blocks:
            if boundary is None or b.boundary == boundary:
                off = b.row0 - self.start
                if tile_ss is None:
                    ops.spmv(b.payload, x, y_local[off:off + b.nrow], **kw)
                else:   # blocks start on tile boundaries of the slice (stencil_row_blocks, from_csr)
                    assert off % 32 == 0
                    ops.spmv(b.payload, x, y_local[off:off + b.nrow], tile_ss[off // 32:off // 32 + (b.nrow + 31) // 32], **kw)

    def needed_ranges(self) -> list[tuple[int, int, int]]:
        """(owner rank, lo, hi) pieces of x outside the own slice that the boundary blocks read."""
        lo = min((b.col_min for b in self.blocks if b.boundary), default=self.start)
        hi = max((b.col_max for b in self.blocks if b.boundary), default=self.start + self.count - 1) + 1
        out = []
        for r in range(self.world):
            if r == self.rank:
                continue
            s, c = partition_rows(self.N, self.world, r)
            a, b = max(lo, s), min(hi, s + c)
            if a < b:
                out.append((r, a, b))
        return out


class PowerIteration:
    def __init__(self, A: PartitionedCSR, ops, exchange: str = "allgather", overlap: bool = True, group=None, seed: int = 11,
                 reserve_sms: int = 16):
        self.A, self.ops, self.exchange, self.overlap, self.group = A, ops, exchange, overlap, group
        self.world, self.rank = A.world, A.rank
        if hasattr(ops, "reserve_sms"):
            for b in A.blocks:   # interior blocks run next to the x refresh; boundary blocks run alone
                # Only the NCCL all-gather needs whole SMs (its CTAs do not fit beside the persistent
                # SpMV CTA); the push / halo / barrier kernels are small enough to co-reside.
                # Measured on 2 GPUs (profiles/r01_power_2gpu.txt): all-gather 5.23 -> 4.43 ms with 16
                # SMs left free, while push/halo lose ~5 % when SMs are taken away.
                free = reserve_sms if (self.world > 1 and overlap and not b.boundary and exchange == "allgather") else 0
                ops.reserve_sms(b.payload, free)
        self.y = ops.empty(A.count)
        self.ss = ops.scalar()
        # sum of squares in the canonical order (csrc/tree_sum.cuh): per-tile partials from the SpMV's epilogue,
        # index-bit tree over the tiles, then over the ranks - the same bits on 1, 2, 4, 8 GPUs for aligned row blocks
        self.tile_ss = ops.empty((A.count + 31) // 32)
        self.rank_ss = ops.empty(self.world) if self.world > 1 else None
        self.symm = None
        if self.world > 1 and exchange in ("fused", "push", "cepush", "halo", "xchg") and hasattr(ops, "symmetric_x"):
            self.x = ops.symmetric_x(A.N)     # CPU test ops: plain memory, exchanges emulated over gloo
            self.peer_ptrs = None
        elif self.world > 1 and exchange in ("fused", "push", "cepush", "halo", "xchg"):
            import torch.distributed._symmetric_memory as symm_mem
            self.x = symm_mem.empty(A.N, dtype=torch.float64, device=ops.device)
            self.symm = symm_mem.rendezvous(self.x, group=dist.group.WORLD.group_name if group is None else group.group_name)
            self.peer_ptrs = [int(self.symm.buffer_ptrs[r]) for r in range(self.world)]
        else:
            self.x = ops.empty(A.N)
            self.peer_ptrs = None
        self.comm_stream = torch.cuda.Stream(device=ops.device) if (self.world > 1 and ops.device.type == "cuda") else None
        self.x_ready = None  # event: remote parts of x are fresh
        self.iter = 0
        self.trace = None    # list of (phase name, event) when tracing (PowerIteration.trace_on)
        if self.world > 1 and exchange == "xchg":
            pg = dist.group.WORLD if group is None else group
            ops.xchg_setup(self.world, self.rank, getattr(pg, "group_name", ""))
            needs = A.needed_ranges()
            all_needs = [None] * self.world
            dist.all_gather_object(all_needs, needs, group=group)
            ops.xchg_dests(pushes_from_needs(all_needs, self.rank), self.peer_ptrs)
            self.src_mask = 0
            for owner, _, _ in needs:
                self.src_mask |= 1 << owner
        self.equal_split = A.N % self.world == 0
        self._init_x(seed)

    def _init_x(self, seed):
        # x0 = the same counter-hash vector on every rank (thsp_gen_vector_f64 / oracle_gen_vector)
        if hasattr(self.ops, "init_x"):
            self.ops.init_x(self.x, seed)
        else:
            H = self.ops.H
            self.ops.check(self.ops.lib.thsp_gen_vector_f64(C.c_int64(self.A.N), C.c_uint64(seed), self.ops.ptr(self.x), self.ops.stream()))
        if self.symm is not None:
            torch.cuda.synchronize()
            self.symm.barrier()

    # ---- one iteration ------------------------------------------------------------------
    def step(self):
        A, ops = self.A, self.ops
        if self.world > 1 and self.exchange == "xchg":
            return self._step_xchg()
        cur = torch.cuda.current_stream() if self.comm_stream is not None else None
        self._mark("start")
        if self.world > 1 and self.overlap:
            A.spmv(ops, self.x, self.y, boundary=False, tile_ss=self.tile_ss)      # needs only the own slice of x
            self._mark("interior rows")
            self._wait_refresh(cur)
            self._mark("wait for refresh")
            A.spmv(ops, self.x, self.y, boundary=True, tile_ss=self.tile_ss)
            self._mark("boundary rows")
        else:
            self._wait_refresh(cur)
            A.spmv(ops, self.x, self.y, tile_ss=self.tile_ss)
            self._mark("all rows")
        ops.tree_sum(self.tile_ss, self.ss)
        self._mark("sum of squares")
        if self.world > 1:
            dist.all_gather_into_tensor(self.rank_ss, self.ss, group=self.group)        # 8 bytes per rank
            ops.tree_sum(self.rank_ss, self.ss)
            self._mark("all-gather of the partials")
        self._scale_and_refresh(cur)
        self._mark("scale (+ refresh launch)")

    # ---- phase trace: where an iteration's time goes (CUDA events on the main stream) -----------
    def trace_on(self):
        self.trace = []

    def _mark(self, name):
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.trace.append((name, e))

    def trace_report(self):
        """Mean milliseconds between consecutive marks, keyed by the phase that ENDS at the mark."""
        torch.cuda.synchronize()
        acc, cnt = {}, {}
        for (_, e0), (name, e1) in zip(self.trace[:-1], self.trace[1:]):
            acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1)
            cnt[name] = cnt.get(name, 0) + 1
        return {k: acc[k] / cnt[k] for k in acc}

    def _step_xchg(self):
        """One stream, no collective: interior rows | wait for the neighbours' pieces of the previous
        step | boundary rows | partial sum published to all ranks | normalise + push the pieces the
        neighbours read."""
        A, ops = self.A, self.ops
        k = self.iter + 1
        self._mark("start")
        if self.overlap:
            A.spmv(ops, self.x, self.y, boundary=False, tile_ss=self.tile_ss)
            self._mark("interior rows")
            ops.wait_halo(k - 1, self.src_mask)
            self._mark("wait for neighbours")
            A.spmv(ops, self.x, self.y, boundary=True, tile_ss=self.tile_ss)
            self._mark("boundary rows")
        else:
            ops.wait_halo(k - 1, self.src_mask)
            A.spmv(ops, self.x, self.y, tile_ss=self.tile_ss)
            self._mark("all rows")
        ops.norm_scale_push(self.y, self.tile_ss, k, self.x, A.start, self.ss)
        self._mark("tree + publish + combine + scale + push")
        self.iter = k

    def _wait_refresh(self, cur):
        if self.x_ready is not None and cur is not None:
            cur.wait_event(self.x_ready)

    def _scale_and_refresh(self, cur):
        A, ops = self.A, self.ops
        if self.world == 1:
            ops.scale_into(self.y, self.ss, self.x, A.start)
            return
        if self.exchange == "fused":
            # One kernel: normalise and store into every replica over NVLink.  The barrier makes
            # the stores of all ranks visible before anyone's boundary rows read them.
            ops.scale_into(self.y, self.ss, self.x, A.start, peer_ptrs=self.peer_ptrs)
            self._barrier_async(cur)
            return
        ops.scale_into(self.y, self.ss, self.x, A.start)
        if self.exchange == "halo":
            self._barrier_async(cur, pull=True)
            return
        if self.exchange == "push":
            self._barrier_async(cur, push=True)
            return
        if self.exchange == "cepush":
            self._barrier_async(cur, ce=True)
            return
        # NCCL all-gather of the slices, in place, on the side stream
        done = torch.cuda.Event() if cur is not None else None
        if cur is not None:
            done.record(cur)
            self.comm_stream.wait_event(done)
        ctx = torch.cuda.stream(self.comm_stream) if self.comm_stream is not None else _Null()
        with ctx:
            own = self.x[A.start:A.start + A.count]
            _all_gather_slices(self.x, own, A.N, self.world, self.equal_split, self.group)
            if self.comm_stream is not None:
                self.x_ready = torch.cuda.Event()
                self.x_ready.record(self.comm_stream)

    def _barrier_async(self, cur, pull=False, push=False, ce=False):
        """Side-stream part of the refresh, so that interior rows can be multiplied meanwhile:
        `push`: store the own (already normalised) slice into every other replica over NVLink with
        one kernel, then a symmetric-memory barrier; `halo`: barrier, then pull the needed pieces
        of x from their owners' replicas; `fused`: just the barrier."""
        done = torch.cuda.Event()
        done.record(cur)
        self.comm_stream.wait_event(done)
        with torch.cuda.stream(self.comm_stream):
            if push:
                A = self.A
                if not hasattr(self, "_one"):
                    self._one = torch.ones(1, dtype=torch.float64, device=self.x.device)
                others = [p for r, p in enumerate(self.peer_ptrs) if r != self.rank]
                own = self.x[A.start:A.start + A.count]
                self.ops.scale_into(own, self._one, self.x, A.start, peer_ptrs=others)   # x * (1/sqrt(1)) == x exactly
            if ce:
                # Every rank sends its slice to every other rank: at step k rank r sends to rank (r + k) mod world, each step on
                # its own stream, so that the world-1 steps are world-1 PERMUTATIONS running side by side - no GPU receives
                # from two senders in one step, every link carries one copy at a time.  (Round 1 issued the copies in rank
                # order on one stream: all ranks wrote to rank 0 first, then to rank 1, ...: 4.7 ms at 8 GPUs.)
                A = self.A
                own = self.x[A.start:A.start + A.count]
                if not hasattr(self, "_ce_streams"):
                    self._ce_streams = [torch.cuda.Stream(device=self.x.device) for _ in range(self.world - 1)]
                for k, st in enumerate(self._ce_streams, start=1):
                    dst = (self.rank + k) % self.world
                    st.wait_event(done)
                    with torch.cuda.stream(st):
                        self.ops.check(self.ops.lib.thsp_memcpy_d2d(C.c_void_p(self.peer_ptrs[dst] + A.start * 8), self.ops.ptr(own),
                                                                    C.c_size_t(A.count * 8), self.ops.stream()))
                        ev = torch.cuda.Event()
                        ev.record(st)
                    self.comm_stream.wait_event(ev)
            self.symm.barrier()
            if pull:
                for owner, lo, hi in self.A.needed_ranges():
                    src = self.symm.get_buffer(owner, (self.A.N,), torch.float64)
                    self.x[lo:hi].copy_(src[lo:hi], non_blocking=True)
                # no second barrier: an owner overwrites its slice only after the next all-reduce,
                # which this rank joins after the rows that read these pieces have been multiplied
            self.x_ready = torch.cuda.Event()
            self.x_ready.record(self.comm_stream)

    def norm(self) -> float:
        """||A x|| of the last step (the power-iteration eigenvalue estimate), on the host."""
        if self.world > 1 and self.exchange == "xchg" and hasattr(self.ops, "xchg_timed_out") and self.ops.xchg_timed_out():
            raise RuntimeError("flag-based exchange: a wait on a peer gave up (peer stalled or died)")
        return math.sqrt(float(self.ss.item()))

    def y_hash(self) -> int:
        """Order-independent fingerprint of this rank's rows of y at their global positions (thsp_hash_f64): the ranks'
        values add up (mod 2^64) to the fingerprint of the whole vector on one GPU."""
        return self.ops.hash(self.y, self.A.start)

    def bytes_per_step(self) -> int:
        """Algorithmic HBM bytes this rank moves per iteration: CSR stream + x once +
        y write (the sums of squares leave the SpMV's epilogue: y is not read for them), y read + x-slice write (scale)."""
        a = self.A
        rows = a.count
        lo = min((b.col_min for b in a.blocks), default=0)
        hi = max((b.col_max for b in a.blocks), default=-1)
        x_read = max(0, hi - lo + 1)   # the columns this rank's rows touch (own slab + halo), not all of x
        return a.nnz_local * 12 + (rows + len(a.blocks)) * 4 + x_read * 8 + rows * 8 + rows * 16


class DeferredPowerIteration:
    """The same loop with the normalisation deferred into the next product.  x = y / ||y|| is never stored: the vector
    stays as the SpMV left it and the next SpMV multiplies every gathered y_j by 1/||y|| on the fly
    (thsp_csr_plan_spmv_scaled_f64) - the product mul_rn(y_j, 1/||y||) is the very number the normalising pass would
    have written, so y and ||y|| have the same bits as in PowerIteration, step by step, on any number of GPUs, while one
    read and one write of the vector per iteration are gone.  Two vectors alternate (a product cannot overwrite what it
    gathers from); a rank's rows of y ARE its slice of the next input.  exchange: "deferred" on one GPU; "xchgd" across
    GPUs = the flag-based exchange of `xchg` moving the RAW pieces (thsp_xchg_norm_push_f64: nothing in the vector half
    waits for the norm except the scalar itself).  `x` materialises the normalised vector on demand."""
    deferred = True

    def __init__(self, A: PartitionedCSR, ops, exchange: str = "xchgd", overlap: bool = True, group=None, seed: int = 11,
                 reserve_sms: int = 16):
        self.A, self.ops, self.exchange, self.overlap, self.group = A, ops, exchange, overlap, group
        self.world, self.rank = A.world, A.rank
        if hasattr(ops, "reserve_sms"):
            for b in A.blocks:
                ops.reserve_sms(b.payload, 0)
        self.ss, self.inv = ops.scalar(), ops.scalar()
        self.inv.fill_(1.0)   # x0 is taken as it is: multiplying by 1.0 changes no bit
        self.tile_ss = ops.empty((A.count + 31) // 32)
        self.symm = []
        if self.world > 1 and hasattr(ops, "symmetric_x"):
            self.buf = [ops.symmetric_x(A.N), ops.symmetric_x(A.N)]     # CPU test ops: plain memory, pushes emulated over gloo
            peer_sets = [None, None]
        elif self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            name = dist.group.WORLD.group_name if group is None else group.group_name
            self.buf, peer_sets = [], []
            for _ in range(2):
                t = symm_mem.empty(A.N, dtype=torch.float64, device=ops.device)
                h = symm_mem.rendezvous(t, group=name)
                self.buf.append(t)
                self.symm.append(h)
                peer_sets.append([int(h.buffer_ptrs[r]) for r in range(self.world)])
        else:
            self.buf = [ops.empty(A.N), ops.empty(A.N)]
        self.cur = 0     # the vector the next product reads
        self.iter = 0
        self.trace = None
        self.src_mask = 0
        if self.world > 1:
            pg = dist.group.WORLD if group is None else group
            ops.xchg_setup(self.world, self.rank, getattr(pg, "group_name", ""))
            needs = A.needed_ranges()
            all_needs = [None] * self.world
            dist.all_gather_object(all_needs, needs, group=group)
            ops.xchg_dests2(pushes_from_needs(all_needs, self.rank), peer_sets)
            for owner, _, _ in needs:
                self.src_mask |= 1 << owner
        if hasattr(ops, "init_x"):
            ops.init_x(self.buf[0], seed)
        else:
            ops.check(ops.lib.thsp_gen_vector_f64(C.c_int64(A.N), C.c_uint64(seed), ops.ptr(self.buf[0]), ops.stream()))
        if self.symm:
            torch.cuda.synchronize()
            self.symm[0].barrier()

    # the phase trace of PowerIteration
    trace_on = PowerIteration.trace_on
    _mark = PowerIteration._mark
    trace_report = PowerIteration.trace_report

    @property
    def y(self):
        """this rank's rows of the last product (they sit in the vector the next product reads)"""
        return self.buf[self.cur][self.A.start:self.A.start + self.A.count]

    @property
    def x(self):
        """the normalised vector y / ||y|| (x0 before the first step): own slice and the pieces this rank reads"""
        return self.ops.scaled(self.buf[self.cur], self.inv)

    def step(self):
        A, ops = self.A, self.ops
        k = self.iter + 1
        xin, xout = self.buf[self.cur], self.buf[1 - self.cur]
        yown = xout[A.start:A.start + A.count]
        self._mark("start")
        if self.world > 1 and self.overlap:
            A.spmv(ops, xin, yown, boundary=False, tile_ss=self.tile_ss, xscale=self.inv)
            self._mark("interior rows")
            ops.wait_halo(k - 1, self.src_mask)
            self._mark("wait for neighbours")
            A.spmv(ops, xin, yown, boundary=True, tile_ss=self.tile_ss, xscale=self.inv)
            self._mark("boundary rows")
        else:
            if self.world > 1:
                ops.wait_halo(k - 1, self.src_mask)
            A.spmv(ops, xin, yown, tile_ss=self.tile_ss, xscale=self.inv)
            self._mark("all rows")
        if self.world > 1:
            ops.norm_push(yown, self.tile_ss, k, xout, A.start, self.ss, self.inv, 1 - self.cur)
            self._mark("tree + publish + push + combine")
        else:
            ops.tree_sum(self.tile_ss, self.ss)
            ops.inv_sqrt(self.ss, self.inv)
            self._mark("sum of squares + 1/sqrt")
        self.cur = 1 - self.cur
        self.iter = k

    def norm(self) -> float:
        if self.world > 1 and hasattr(self.ops, "xchg_timed_out") and self.ops.xchg_timed_out():
            raise RuntimeError("flag-based exchange: a wait on a peer gave up (peer stalled or died)")
        return math.sqrt(float(self.ss.item()))

    def y_hash(self) -> int:
        return self.ops.hash(self.y, self.A.start)

    def bytes_per_step(self) -> int:
        """CSR stream + the columns read once + y written; no normalising pass"""
        a = self.A
        lo = min((b.col_min for b in a.blocks), default=0)
        hi = max((b.col_max for b in a.blocks), default=-1)
        return a.nnz_local * 12 + (a.count + len(a.blocks)) * 4 + max(0, hi - lo + 1) * 8 + a.count * 8


DEFERRED_MODES = ("deferred", "xchgd")


def make_iteration(A, ops, exchange="allgather", **kw):
    """PowerIteration, or DeferredPowerIteration for the modes that fold the normalisation into the next product."""
    cls = DeferredPowerIteration if exchange in DEFERRED_MODES else PowerIteration
    return cls(A, ops, exchange=exchange, **kw)


def _all_gather_slices(x, own, n, world, equal_split, group):
    """Every rank's slice into every replica of x.  Equal slices: one in-place all-gather.  The
    reference's rule gives the last block the remainder (src/mat_vec.cpp:245-246); collectives
    want equal counts, so an uneven split falls back to one broadcast per owner."""
    if equal_split:
        dist.all_gather_into_tensor(x, own, group=group)
        return
    for r in range(world):
        s, c = partition_rows(n, world, r)
        dist.broadcast(x[s:s + c], src=r, group=group)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# =========================================================================================
class ColumnPartitionedCSC:
    """CSC across the GPUs: what CSCMatrixMatVectorNuma does with threads
    (src/mat_vec.cpp:299-366 - equal COLUMN blocks, column pointers rebased, a full-length private y per
    block) plus the step the reference leaves out: the private y's are never added up there.  Here every
    rank multiplies its column block by its slice of x into a full-length partial y, and one
    reduce-scatter (NCCL over NVLink) leaves each rank with its slice of the sum; ranks own equal row and
    column slices except the last, which takes the remainder (:233,245-246) - an uneven split falls back to
    an all-reduce."""

    def __init__(self, nrow: int, ncol: int, col_ptr, row_ind, values, rank: int, world: int, ops, group=None):
        self.nrow, self.ncol, self.rank, self.world, self.ops, self.group = nrow, ncol, rank, world, ops, group
        self.c0, self.ncol_local = partition_rows(ncol, world, rank)
        self.r0, self.nrow_local = partition_rows(nrow, world, rank)
        e0, e1 = int(col_ptr[self.c0]), int(col_ptr[self.c0 + self.ncol_local])
        sub_cp = (col_ptr[self.c0:self.c0 + self.ncol_local + 1] - e0).to(torch.int32).contiguous()   # rebased, :331-334
        self.block = ops.csc_block(nrow, self.ncol_local, sub_cp, row_ind[e0:e1].contiguous(), values[e0:e1].contiguous())
        self.partial = ops.empty(nrow)
        self.equal_rows = nrow % world == 0

    def spmv(self, x_slice, y_slice):
        """y_slice = (A x)[own rows]; x_slice = x[own columns]."""
        self.partial.zero_()
        self.ops.csc_spmv(self.block, x_slice, self.partial)     # partial += A[:, own columns] x_slice
        if self.world == 1:
            y_slice.copy_(self.partial)
        elif self.equal_rows and self.partial.is_cuda:
            dist.reduce_scatter_tensor(y_slice, self.partial, group=self.group)
        else:
            dist.all_reduce(self.partial, group=self.group)
            y_slice.copy_(self.partial[self.r0:self.r0 + self.nrow_local])


# =========================================================================================
def e2e_spmv_step(it: "PowerIteration", xh, yh):
    """The distributed SpMV as a caller with HOST buffers sees it: this rank's slice of x comes
    from pinned host memory, the replicas are refreshed (NCCL all-gather), y = A x, and this
    rank's slice of y goes back to pinned host memory.  Returns after the copy has landed."""
    A = it.A
    own = it.x[A.start:A.start + A.count]
    own.copy_(xh, non_blocking=True)
    if it.world > 1:
        _all_gather_slices(it.x, own, A.N, it.world, it.equal_split, it.group)
    A.spmv(it.ops, it.x, it.y)
    yh.copy_(it.y, non_blocking=True)
    torch.cuda.current_stream().synchronize()


def _all_ok(ok: bool, world: int, group, device) -> bool:
    """True on every rank only if `ok` on every rank (also a barrier): ranks must leave a failed step together."""
    if world <= 1:
        return ok
    t = torch.tensor([0.0 if ok else 1.0], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item()) == 0.0


class SharedHostVector:
    """One full-length fp64 vector in host memory that every process of the box maps (a /dev/shm segment) and
    page-locks (thsp_host_register) - where the reference keeps x: `CSRMatrixMatVectorNuma` hands every NUMA thread
    the caller's whole host vector and each copies what it needs (src/mat_vec.cpp:257,266).  Rank 0 creates the
    segment; call from all ranks.  Raises OSError on EVERY rank if any rank could not map or lock it."""

    def __init__(self, n: int, rank: int, world: int, device, group=None, tag: str = "x"):
        import numpy as np
        from .lib import load
        self.lib = load()
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.path = f"/dev/shm/thsp_{tag}_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
        self.tensor = self.array = None
        self.registered = False
        ok = True
        if rank == 0:
            try:
                with open(self.path, "wb") as f:
                    # reserve the pages now: a /dev/shm smaller than the vector must fail here (ENOSPC), not with a
                    # bus error when the vector is first written
                    os.posix_fallocate(f.fileno(), 0, max(n * 8, 1))
            except OSError:
                ok = False
        if not _all_ok(ok, world, group, device):
            self._unlink()
            raise OSError(f"cannot create {self.path}")
        try:
            self.array = np.memmap(self.path, dtype=np.float64, mode="r+", shape=(n,))
            self.tensor = torch.from_numpy(self.array)
            self.registered = self.lib.thsp_host_register(C.c_void_p(self.tensor.data_ptr()), C.c_size_t(n * 8)) == 0
            ok = self.registered
        except (OSError, ValueError):
            ok = False
        if not _all_ok(ok, world, group, device):
            self.close()
            raise OSError(f"cannot map or page-lock {self.path} on every rank")

    def _unlink(self):
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass

    def close(self):
        if self.registered:
            self.lib.thsp_host_unregister(C.c_void_p(self.tensor.data_ptr()))
            self.registered = False
        _all_ok(True, self.world, self.group, self.device)   # nobody unlinks while another rank is still mapping
        self.tensor = self.array = None
        self._unlink()


def e2e_spmv_step_shared(it: "PowerIteration", x_host, yh):
    """The same call when x sits in ONE host vector all ranks can read (SharedHostVector): each rank pulls the
    window of x its row blocks read straight from there - own slice plus the neighbours' planes - chunk by chunk
    through thsp_csr_plan_spmv_host_f64, multiplies a chunk as soon as its window has arrived and sends its rows of
    y back while the next chunk is coming in.  No collective: the refresh of the replicas is the upload itself."""
    A = it.A
    ops = it.ops
    for b in A.blocks:
        off = b.row0 - A.start
        ops.check(ops.lib.thsp_csr_plan_spmv_host_f64(b.payload.plan(), C.c_void_p(x_host.data_ptr()),
                                                      C.c_void_p(yh.data_ptr() + off * 8), ops.ptr(it.x),
                                                      C.c_void_p(it.y.data_ptr() + off * 8), 0, ops.stream()))


def measure(n, rank, world, device, steps, warmup, modes, overlap, ClockSampler, with_e2e=True, reserve_sms=16, hash_parts=0):
    """Time `steps` power-iteration steps per x-refresh mode (device events, max over ranks).  Every mode starts from
    the same x0 and runs max(3, warmup) + steps steps, so `norm` and the fingerprint of y ("y_hash": this rank's rows;
    "y_hash_parts": with hash_parts = G on one GPU, the rows each of G ranks would own) can be compared across GPU counts."""
    from .lib import launch_count
    ops = CudaOps(device)
    A = PartitionedCSR.stencil27(n, rank, world, ops)
    results = {}
    sync_token = torch.zeros(1, device=device)
    for mode in modes:
        try:
            it = make_iteration(A, ops, exchange=mode, overlap=overlap, reserve_sms=reserve_sms)
        except Exception as e:  # e.g. symmetric memory not available on this box
            results[mode] = {"error": f"{type(e).__name__}: {e}"[:300]}
            continue
        for _ in range(max(3, warmup)):
            it.step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = launch_count()
        with ClockSampler(device.index or 0) as clk:
            if world > 1:
                # The host threads leave the barrier above milliseconds apart; a collective queued on the
                # stream right before the first event lines the GPUs up, so the timed region holds the
                # `steps` iterations and not the launch skew of the slowest host thread.
                dist.all_reduce(sync_token)
            e0.record()
            for _ in range(steps):
                it.step()
            e1.record()
            torch.cuda.synchronize()
        launches = launch_count() - l0
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=device)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        results[mode] = {"ms_per_step": float(ms.item()), "norm": it.norm(), "launches": int(launches), "clocks": clk.summary(),
                         "bytes_per_step_rank": it.bytes_per_step(), "y_hash": it.y_hash(), "steps_done": max(3, warmup) + steps}
        if hash_parts:
            results[mode]["y_hash_parts"] = [ops.hash(it.y[s0:s0 + c0], s0) for s0, c0 in (partition_rows(A.N, hash_parts, r) for r in range(hash_parts))]
        if os.environ.get("THSP_POWER_TRACE"):   # where the time goes, rank by rank (a separate, untimed run)
            it.trace_on()
            for _ in range(10):
                it.step()
            rep = it.trace_report()
            results[mode]["phases_ms"] = {k: round(v, 4) for k, v in rep.items()}
            print(f"[trace rank {rank} {mode}] " + "  ".join(f"{k}: {v:.3f}" for k, v in rep.items()), flush=True)
            it.trace = None
        if with_e2e and "e2e" not in results and not getattr(it, "deferred", False):
            xh = torch.empty(A.count, dtype=torch.float64).pin_memory()
            yh = torch.empty(A.count, dtype=torch.float64).pin_memory()
            xh.copy_(it.x[A.start:A.start + A.count].cpu())
            for _ in range(2):
                e2e_spmv_step(it, xh, yh)
            if world > 1:
                dist.barrier()
            k = max(3, min(steps, 10))
            t0 = time.perf_counter()
            for _ in range(k):
                e2e_spmv_step(it, xh, yh)
            dt = torch.tensor([(time.perf_counter() - t0) / k], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            results["e2e"] = {"ms_per_step": float(dt.item()) * 1e3, "h2d_bytes_per_step": A.count * 8 * world,
                              "d2h_bytes_per_step": A.count * 8 * world,
                              "call": "distributed y = A x: pinned x slices in, NCCL all-gather, SpMV, pinned y slices out"}
            # the same with x in one host vector shared by the ranks (where the reference keeps it): windows pulled from
            # there and pipelined against the kernels and the download of y; kept if faster
            xs = None
            try:
                want = yh.clone()
                xs = SharedHostVector(A.N, rank, world, device, it.group)
                if rank == 0:
                    xs.tensor.copy_(it.x.cpu())
                ok = _all_ok(True, world, it.group, device)
                try:
                    for _ in range(3):   # eager, graph capture, graph replay
                        e2e_spmv_step_shared(it, xs.tensor, yh)
                except RuntimeError as exc:
                    ok = False
                    print(f"[rank {rank}] shared-x e2e failed: {exc}", flush=True)
                ok = _all_ok(ok and torch.equal(yh, want), world, it.group, device)
                if ok:
                    t0 = time.perf_counter()
                    for _ in range(k):
                        e2e_spmv_step_shared(it, xs.tensor, yh)
                    dt2 = torch.tensor([(time.perf_counter() - t0) / k], dtype=torch.float64, device=device)
                    up = torch.tensor([float(sum(b.col_max - b.col_min + 1 for b in A.blocks) * 8)], dtype=torch.float64, device=device)
                    if world > 1:
                        dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
                        dist.all_reduce(up, op=dist.ReduceOp.SUM)
                    shared = {"ms_per_step": float(dt2.item()) * 1e3, "h2d_bytes_per_step": int(up.item()),
                              "d2h_bytes_per_step": A.count * 8 * world, "bit_identical_to_allgather_path": True,
                              "call": "distributed y = A x: every rank pulls its window of x from one pinned host vector shared by the "
                                      "ranks (thsp_csr_plan_spmv_host_f64 per row block: upload, SpMV and download of y pipelined), "
                                      "no collective"}
                    results["e2e_allgather"] = results["e2e"]
                    if shared["ms_per_step"] < results["e2e"]["ms_per_step"]:
                        results["e2e"] = shared
                    else:
                        results["e2e_shared_x"] = shared
                else:
                    results["e2e_shared_x"] = {"unavailable": "the shared-x path failed or disagreed with the all-gather path on some rank"}
            except OSError as exc:   # no /dev/shm, registration refused (raised on every rank): keep the all-gather number
                results["e2e_shared_x"] = {"unavailable": str(exc)[:200]}
            finally:
                if xs is not None and xs.tensor is not None:
                    xs.close()
        del it
        torch.cuda.empty_cache()
    return A, results


MODE_NOTES = {
    "xchgd": "normalisation deferred into the next SpMV (every gathered y_j times 1/||y|| inside the stream kernel: same bits, no "
             "pass over the vector); the raw pieces the neighbours read and the partial sums of squares go straight into their memory "
             "over NVLink with flags (thsp_xchg_norm_push_f64); no collective call in the loop",
    "deferred": "one GPU: normalisation deferred into the next SpMV, no pass over the vector",
    "xchg": "each rank stores the pieces of x its neighbours read, and its partial sum of squares, straight into their memory "
            "over NVLink and raises a flag (csrc/exchange.cu); no collective call in the loop",
    "allgather": "NCCL in-place all-gather of every slice into every replica (BASELINE.json's wording), overlapped with the interior rows",
    "halo": "NCCL all-reduce of the scalar, then the needed pieces of x are pulled from their owners (peer copies)",
    "cepush": "own slice copied into every replica by the copy engines (peer cudaMemcpyAsync), overlapped with the interior rows",
    "push": "own slice stored into every replica by one kernel over NVLink peer pointers, on a side stream",
    "fused": "the normalising kernel stores into all replicas itself",
}


def bench_main(args, METRIC, UNIT, csr_bytes, peak_hbm, ClockSampler, extra=None):
    """bench.py --gpus N under torchrun (one rank per GPU).  Returns non-zero when the N-GPU run does not reproduce the
    one-GPU run (parity block of the line)."""
    rc = 0
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    n = args.grid
    if args.reserve_sms < 0:
        # NCCL's all-gather wants more CTAs the more peers it talks to: measured best 16 SMs at 2 GPUs,
        # 48 at 8 (profiles/r01_power_8gpu.txt)
        args.reserve_sms = 16 * max(1, int(math.log2(max(world, 2))))
    modes = [m for m in args.exchange.split(",") if m] if world > 1 else ["allgather"]
    # The same workload on ONE GPU first (rank 0; 512^3 is 45 GB of 180): the denominator of the speed-up, and the norm
    # and the fingerprints of y that the N-GPU run must reproduce bit for bit (same x0, same number of steps).
    one = None
    if world > 1 and not getattr(args, "no_one_gpu", False):
        if rank == 0:
            try:
                A1, res1 = measure(n, 0, 1, device, args.steps, args.warmup, ["deferred", "allgather"], True, ClockSampler, with_e2e=False,
                                   hash_parts=world)
                eager, lazy = res1["allgather"], res1.get("deferred", {})
                # the denominator is the faster one-GPU loop; both must agree bit for bit (same products, same sums)
                use_lazy = "ms_per_step" in lazy and lazy["ms_per_step"] < eager["ms_per_step"]
                one = dict(lazy if use_lazy else eager)
                one["loop"] = "normalisation deferred into the next SpMV" if use_lazy else "SpMV, sum of squares, normalising pass"
                one["eager_ms_per_step"] = eager["ms_per_step"]
                if "ms_per_step" in lazy:
                    one["deferred_ms_per_step"] = lazy["ms_per_step"]
                    one["deferred_equals_eager_bits"] = bool(lazy["norm"] == eager["norm"] and lazy["y_hash_parts"] == eager["y_hash_parts"])
                else:
                    one["deferred"] = lazy
                one["row_blocks"] = len(A1.blocks)
                del A1, res1
            except Exception as e:   # e.g. out of memory on a smaller GPU: the line then says so instead of a ratio
                one = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
        dist.barrier()
    A, results = measure(n, rank, world, device, args.steps, args.warmup, modes, not args.no_overlap, ClockSampler,
                         reserve_sms=args.reserve_sms)
    if world > 1:   # every rank's fingerprint of its rows of y, per mode, to rank 0
        mine = {m: v.get("y_hash") for m, v in results.items() if isinstance(v, dict) and "y_hash" in v}
        allh = [None] * world
        dist.all_gather_object(allh, mine)
    if rank == 0:
        primary = next(m for m in modes if "ms_per_step" in results.get(m, {}))
        nnz_total = (3 * n - 2) ** 3
        flops = 2.0 * nnz_total
        r = results[primary]
        ms = r["ms_per_step"]
        peak, peak_kind = peak_hbm()
        achieved = r["bytes_per_step_rank"] / (ms * 1e-3) / 1e9
        e2e = results.get("e2e")
        line = {
            "metric": METRIC, "value": round(flops / (ms * 1e-3) / 1e9, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms, 5), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "row-partitioned fp64 CSR power iteration (SpMV + sum of squares over all ranks + normalise + x refresh), "
                                   f"27-point stencil {n}^3 generated on device (BASELINE configs[4])",
                       "rows": n ** 3, "nnz": nnz_total, "x_refresh": primary, "x_refresh_note": MODE_NOTES.get(primary, ""), "overlap": not args.no_overlap, "sms_reserved_for_refresh": args.reserve_sms,
                       "partition": f"equal row blocks x{world} (src/mat_vec.cpp:233)", "row_blocks_rank0": len(A.blocks),
                       "cache": "per-rank inputs larger than L2 (126 MB)", "norm": r["norm"]},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": None, "peak_source": peak_kind,
                         "note": "rank 0's algorithmic HBM bytes of one whole iteration / iteration time"},
            "e2e": ({"value": round(flops / (e2e["ms_per_step"] * 1e-3) / 1e9, 2), "unit": UNIT, "ms_per_step": round(e2e["ms_per_step"], 4),
                     "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                     "call": e2e.get("call", "")} if e2e else None),
            "gpu_launches": r["launches"], "clocks": r["clocks"],
            "x_refresh_modes": {m: ({"ms_per_step": round(v["ms_per_step"], 5), "gflops": round(flops / (v["ms_per_step"] * 1e-3) / 1e9, 2),
                                     "norm": v["norm"]} if "ms_per_step" in v else v) for m, v in results.items()
                                if m not in ("e2e", "e2e_allgather", "e2e_shared_x")},
        }
        if one is not None and "ms_per_step" in one:
            ms1 = one["ms_per_step"]
            line["same_workload_1gpu"] = {"ms_per_step": round(ms1, 5), "gflops": round(flops / (ms1 * 1e-3) / 1e9, 2), "norm": one["norm"],
                                          "steps_done": one["steps_done"], "row_blocks": one["row_blocks"], "loop": one.get("loop"),
                                          "eager_ms_per_step": round(one["eager_ms_per_step"], 5),
                                          "deferred_ms_per_step": round(one["deferred_ms_per_step"], 5) if "deferred_ms_per_step" in one else None,
                                          "deferred_equals_eager_bits": one.get("deferred_equals_eager_bits"),
                                          "note": "the same power iteration on rank 0's GPU alone (the faster of its two forms), timed in this run before the ranks start"}
            line["speedup_vs_1gpu_same_workload"] = round(ms1 / ms, 4)
            line["efficiency_same_workload"] = round(ms1 / ms / world, 4)
            for m, v in line["x_refresh_modes"].items():
                if "ms_per_step" in v:
                    v["speedup_vs_1gpu_same_workload"] = round(ms1 / v["ms_per_step"], 4)
            # parity across GPU counts: the canonical sum of squares (csrc/tree_sum.cuh) makes the whole loop reproducible
            par = {"steps_compared_after": one["steps_done"]}
            for m in modes:
                v = results.get(m, {})
                if "norm" not in v:
                    continue
                hashes = [h.get(m) for h in allh]
                par[m] = {"norm_bits_equal_1gpu": bool(v["norm"] == one["norm"]),
                          "norm_rel_diff": abs(v["norm"] - one["norm"]) / abs(one["norm"]),
                          "y_hash_equal_1gpu": bool(hashes == one["y_hash_parts"]),
                          "ranks_with_equal_y": int(sum(1 for a, b in zip(hashes, one["y_hash_parts"]) if a == b))}
            line["parity"] = par
        elif one is not None:
            line["same_workload_1gpu"] = one
        for k2 in ("e2e_allgather", "e2e_shared_x"):   # the e2e variant that was not kept as the headline, for comparison
            v = results.get(k2)
            if v and "ms_per_step" in v:
                line[k2] = {"value": round(flops / (v["ms_per_step"] * 1e-3) / 1e9, 2), "unit": UNIT, "ms_per_step": round(v["ms_per_step"], 4),
                            "h2d_bytes_per_step": v["h2d_bytes_per_step"], "d2h_bytes_per_step": v["d2h_bytes_per_step"], "call": v.get("call", "")}
            elif v:
                line[k2] = v
        if extra:
            line.update(extra)
        print(json.dumps(line), flush=True)
        par = line.get("parity", {}).get(primary)
        if par and not (par["norm_rel_diff"] <= 1e-12):
            print(f"bench.py: the {world}-GPU norm differs from the one-GPU norm: {par}", flush=True)
            rc = 1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc
