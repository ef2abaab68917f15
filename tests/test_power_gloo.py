"""Host-side logic of the row-partitioned power iteration on CPU: world_size 2 over gloo.

The package has no CPU arithmetic; this test injects an `ops` object built on the oracle (test
infrastructure) so that partitioning, the x refresh, the scalar all-reduce and the interior /
boundary ordering run exactly as they do under torchrun on GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleOps:
    device = torch.device("cpu")

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        pyoracle.build()
        self.O = pyoracle.Oracle()

    def empty(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def scalar(self):
        return torch.zeros(1, dtype=torch.float64)

    def init_x(self, x, seed):
        x.copy_(torch.from_numpy(self.O.gen_vector(x.numel(), seed)))

    def stencil_block(self, n, r0, r1):
        rp, ci, va = self.O.gen_stencil27_csr(n, r0, r1)
        return (r1 - r0, n ** 3, rp, ci, va), int(rp[-1])

    def csr_block(self, nrow, ncol, row_ptr, col_ind, values):
        return (nrow, ncol, row_ptr.numpy(), col_ind.numpy(), values.numpy()), int(col_ind.numel())

    def spmv(self, payload, x, y, tile_ss=None, xscale=None):
        nrow, ncol, rp, ci, va = payload
        xv = x.numpy()
        if xscale is not None:   # csr_stream_kernel<kScale>: every gathered x_j times the scalar, one rounding
            xv = xv * float(xscale[0])
        y.copy_(torch.from_numpy(self.O.csr_spmv(nrow, ncol, rp, ci, va, xv, np.zeros(nrow))))
        if tile_ss is not None:   # the SpMV's epilogue: sum of squares per tile of 32 rows (csrc/tree_sum.cuh)
            tile_ss.copy_(torch.from_numpy(self.O.tile_sumsq(y.numpy())))

    def tree_sum(self, vals, out):
        out[0] = self.O.tree_sum(vals.numpy())

    def inv_sqrt(self, ss, inv):
        inv[0] = 1.0 / np.sqrt(float(ss[0]))

    def scaled(self, x, scale):
        return torch.from_numpy(x.numpy() * float(scale[0]))

    def hash(self, v, first):
        return self.O.hash_f64(v.numpy(), first)

    def sumsq(self, y, out):
        out[0] = self.O.dot(y.numpy(), y.numpy())

    def scale_into(self, y, sumsq, x, offset, peer_ptrs=None):
        # vec_axpby(1/nrm, y, 0, y, x_slice): the beta == 0 branch (src/vec_vec.cpp:46-53)
        w = self.O.axpby(1.0 / np.sqrt(float(sumsq[0])), y.numpy(), 0.0, y.numpy())
        x[offset:offset + y.numel()] = torch.from_numpy(w)

    def col_range(self, payload):
        ci = payload[3]
        return (int(ci.min()), int(ci.max())) if len(ci) else (0, -1)

    def csc_block(self, nrow, ncol, col_ptr, row_ind, values):
        return (nrow, ncol, col_ptr.numpy(), row_ind.numpy(), values.numpy())

    def csc_spmv(self, payload, x, y):
        nrow, ncol, cp, ri, va = payload
        y.copy_(torch.from_numpy(self.O.csc_spmv(nrow, ncol, cp, ri, va, x.numpy(), y.numpy())))

    # ---- the flag-based exchange (csrc/exchange.cu), emulated with gloo collectives: what is
    # tested here is the host logic - who pushes which piece of x to whom, and in which order
    def symmetric_x(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def xchg_setup(self, world, rank, group_name):
        self.xw, self.xr = world, rank

    def xchg_dests(self, dests, x_peer_ptrs):
        self.dests = dests

    def wait_halo(self, it, src_mask):
        pass   # the emulated push below is synchronous

    def xchg_dests2(self, dests, peer_ptr_sets):
        self.dests = dests

    def norm_push(self, y, tile_ss, it, xout, offset, ss, inv, which):
        """thsp_xchg_norm_push_f64: the sum over all ranks and its 1/sqrt; the RAW pieces into the readers' copies"""
        parts = [None] * self.xw
        dist.all_gather_object(parts, self.O.tree_sum(tile_ss.numpy()))
        tot = self.O.tree_sum(np.array(parts))
        ss[0] = tot
        inv[0] = 1.0 / np.sqrt(tot)
        msgs = [(r, lo, hi, xout[lo:hi].clone()) for r, lo, hi in self.dests]
        for m in msgs:
            assert offset <= m[1] and m[2] <= offset + y.numel(), "a rank may only push pieces of its own slice"
        allm = [None] * self.xw
        dist.all_gather_object(allm, msgs)
        for src, ms in enumerate(allm):
            for r, lo, hi, data in ms:
                if r == self.xr:
                    xout[lo:hi] = data

    def norm_scale_push(self, y, tile_ss, it, x, offset, ss):
        parts = [None] * self.xw
        dist.all_gather_object(parts, self.O.tree_sum(tile_ss.numpy()))
        tot = self.O.tree_sum(np.array(parts))   # tree over rank numbers, like xchg_norm_scale_push_kernel
        ss[0] = tot
        w = torch.from_numpy(self.O.axpby(1.0 / np.sqrt(tot), y.numpy(), 0.0, y.numpy()))
        x[offset:offset + y.numel()] = w
        msgs = [(r, lo, hi, x[lo:hi].clone()) for r, lo, hi in self.dests]
        for lo_hi in msgs:
            assert offset <= lo_hi[1] and lo_hi[2] <= offset + y.numel(), "a rank may only push pieces of its own slice"
        allm = [None] * self.xw
        dist.all_gather_object(allm, msgs)
        for src, ms in enumerate(allm):
            for r, lo, hi, data in ms:
                if r == self.xr:
                    x[lo:hi] = data


def serial_power_iteration(O, n, steps, seed, canonical=False):
    """The loop composed from the reference's calls; canonical: vec_dot(y, y) in the fixed order of
    csrc/tree_sum.cuh instead of the serial sum - what one GPU computes."""
    N = n ** 3
    rp, ci, va = O.gen_stencil27_csr(n)
    x = O.gen_vector(N, seed)
    nrm = 0.0
    for _ in range(steps):
        y = O.csr_spmv(N, N, rp, ci, va, x, np.zeros(N))
        nrm = np.sqrt(O.tree_sum(O.tile_sumsq(y)) if canonical else O.dot(y, y))
        x = O.axpby(1.0 / nrm, y, 0.0, y)
    return x, nrm, y


def _worker(rank, world, port, n, steps, mode, overlap, generic, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arm_spmv_b200 import power
        ops = OracleOps()
        if generic:
            rp, ci, va = ops.O.gen_stencil27_csr(n)
            A = power.PartitionedCSR.from_csr(n ** 3, torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(va), rank, world, ops)
        else:
            A = power.PartitionedCSR.stencil27(n, rank, world, ops, max_block_rows=50)
        it = power.make_iteration(A, ops, exchange=mode, overlap=overlap, seed=5)
        for _ in range(steps):
            it.step()
        # after the last refresh every replica must hold the same, complete x (xchg: only the pieces it reads)
        torch.save({"x": it.x.clone(), "norm": it.norm(), "y": it.y.clone(), "y_hash": it.y_hash(), "start": A.start, "count": A.count, "needs": A.needed_ranges(),
                    "blocks": [(b.row0, b.nrow, b.boundary) for b in A.blocks]}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,overlap,generic,mode", [(2, 6, True, False, "allgather"), (2, 7, False, False, "allgather"),
                                                          (3, 5, True, False, "allgather"), (2, 6, True, True, "allgather"),
                                                          (2, 6, True, False, "xchg"), (3, 5, True, False, "xchg"),
                                                          (3, 6, False, True, "xchg"),
                                                          (2, 8, True, False, "xchg"), (4, 8, True, False, "allgather"),
                                                          (2, 8, False, True, "allgather"),
                                                          (2, 8, True, False, "xchgd"), (3, 5, True, False, "xchgd"),
                                                          (2, 6, False, True, "xchgd"), (4, 8, True, False, "xchgd")])
def test_partitioned_power_iteration_matches_serial(oracle, tmp_path, world, n, overlap, generic, mode):
    steps = 4
    port = 29500 + (os.getpid() + world * 7 + n + len(mode)) % 400
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, port, n, steps, mode, overlap, generic, out), nprocs=world, join=True)
    x_ref, nrm_ref, _ = serial_power_iteration(oracle, n, steps, 5)
    # Row blocks that are whole subtrees of the canonical sum (n^3 a multiple of 32 * world, world a power of two):
    # norm, x and y must equal the ONE-rank loop bit for bit, and the ranks' fingerprints of y add up to its fingerprint.
    aligned = (n ** 3) % (32 * world) == 0 and world & (world - 1) == 0
    if aligned:
        x_one, nrm_one, y_one = serial_power_iteration(oracle, n, steps, 5, canonical=True)
        hsum = 0
    covered = 0
    for r in range(world):
        res = torch.load(f"{out}.{r}")
        assert res["start"] == covered
        if aligned:
            lo, hi = res["start"], res["start"] + res["count"]
            assert res["norm"] == nrm_one, (res["norm"], nrm_one)
            assert res["x"].numpy()[lo:hi].tobytes() == x_one[lo:hi].tobytes()
            assert res["y"].numpy().tobytes() == y_one[lo:hi].tobytes()
            assert res["y_hash"] == oracle.hash_f64(y_one[lo:hi], lo)
            hsum = (hsum + res["y_hash"]) % (1 << 64)
            if r == world - 1:
                assert hsum == oracle.hash_f64(y_one, 0)
        covered += res["count"]
        # rows are multiplied in the reference's order, so only the norm (a sum over ranks) differs in rounding
        if mode in ("xchg", "xchgd"):   # a replica is refreshed only where this rank reads it: own slice + needed pieces
            xs = res["x"].numpy()
            for lo, hi in [(res["start"], res["start"] + res["count"])] + [(a, b) for _, a, b in res["needs"]]:
                assert np.max(np.abs(xs[lo:hi] - x_ref[lo:hi])) <= 1e-13
            assert abs(res["norm"] - nrm_ref) <= 1e-12 * nrm_ref
            continue
        assert np.max(np.abs(res["x"].numpy() - x_ref)) <= 1e-13
        assert abs(res["norm"] - nrm_ref) <= 1e-12 * nrm_ref
        rows = sum(b[1] for b in res["blocks"])
        assert rows == res["count"]
    assert covered == n ** 3


def test_stencil_row_blocks_cover_and_flag_boundaries(oracle):
    sys.path.insert(0, ROOT)
    from arm_spmv_b200 import power
    for n, world in [(6, 2), (5, 3), (8, 4), (4, 4), (9, 8)]:
        N = n ** 3
        rp, ci, _ = oracle.gen_stencil27_csr(n)
        for rank in range(world):
            start, count = power.partition_rows(N, world, rank)
            pieces = power.stencil_row_blocks(n, start, count, world, max_rows=40)
            at = start
            for r0, r1, bnd in sorted(pieces):
                assert r0 == at and r1 > r0
                assert (r0 - start) % 32 == 0, "pieces start on tile boundaries of the slice (canonical sum of squares)"
                at = r1
                cols = ci[rp[r0]:rp[r1]]
                needs_remote = cols.min() < start or cols.max() >= start + count
                assert bnd or not needs_remote, (n, world, rank, r0, r1)   # interior pieces never read remote x
            assert at == start + count


def _csc_worker(rank, world, port, nrow, ncol, nnz, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arm_spmv_b200 import power
        ops = OracleOps()
        ri, cj, v = ops.O.gen_uniform_coo(nrow, ncol, nnz, 7)
        cp, rind, va = ops.O.coo2csc(nrow, ncol, ri, cj, v)
        A = power.ColumnPartitionedCSC(nrow, ncol, torch.from_numpy(cp), torch.from_numpy(rind), torch.from_numpy(va), rank, world, ops)
        x = torch.from_numpy(ops.O.gen_vector(ncol, 2))
        y = torch.zeros(A.nrow_local, dtype=torch.float64)
        A.spmv(x[A.c0:A.c0 + A.ncol_local].clone(), y)
        torch.save({"y": y, "r0": A.r0}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nrow,ncol", [(2, 300, 200), (3, 100, 301)])
def test_column_partitioned_csc_matches_serial(oracle, tmp_path, world, nrow, ncol):
    nnz = 4000
    port = 29900 + (os.getpid() + world + nrow) % 90
    out = str(tmp_path / "csc")
    mp.spawn(_csc_worker, args=(world, port, nrow, ncol, nnz, out), nprocs=world, join=True)
    ri, cj, v = oracle.gen_uniform_coo(nrow, ncol, nnz, 7)
    x = oracle.gen_vector(ncol, 2)
    ref = oracle.coo_spmv(nrow, ncol, ri, cj, v, x, np.zeros(nrow))
    scale = np.bincount(ri, weights=np.abs(v * x[cj]), minlength=nrow)
    got = np.concatenate([torch.load(f"{out}.{r}")["y"].numpy() for r in range(world)])
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.maximum(scale, 1e-300)) <= 1e-12


def test_pushes_are_the_transpose_of_needs():
    sys.path.insert(0, ROOT)
    from arm_spmv_b200 import power
    needs = [[(1, 10, 14)], [(0, 6, 10), (2, 20, 23)], [(1, 15, 20), (1, 11, 12)]]
    assert power.pushes_from_needs(needs, 0) == [(1, 6, 10)]
    assert power.pushes_from_needs(needs, 1) == [(0, 10, 14), (2, 11, 20)]   # one bounding range per reader
    assert power.pushes_from_needs(needs, 2) == [(1, 20, 23)]


def test_partition_matches_reference_rule(oracle):
    sys.path.insert(0, ROOT)
    from arm_spmv_b200 import power
    for n, parts in [(10, 3), (64, 8), (7, 7), (5, 8)]:
        for p in range(parts):
            assert power.partition_rows(n, parts, p) == oracle.partition(n, parts, p)


def _shared_vector_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arm_spmv_b200 import power
        dev = torch.device("cpu")
        assert power._all_ok(True, world, None, dev) is True
        assert power._all_ok(rank != 1, world, None, dev) is False      # one rank's failure is everybody's
        result = "mapped"
        try:
            xs = power.SharedHostVector(1000, rank, world, dev, tag="gloo_test")
            xs.close()
        except OSError:   # no GPU here: page-locking fails - on every rank, after the same collectives
            result = "OSError"
        leftovers = [f for f in os.listdir("/dev/shm") if f.startswith("thsp_gloo_test_")] if os.path.isdir("/dev/shm") else []
        dist.barrier()
        with open(os.path.join(out, f"r{rank}.txt"), "w") as f:
            f.write(result + " " + str(len(leftovers)))
    finally:
        dist.destroy_process_group()


def test_shared_host_vector_fails_on_all_ranks_together(tmp_path):
    """power.SharedHostVector (x of the multi-GPU host-buffer call in one /dev/shm segment): when a rank cannot map or
    page-lock it - as here, without a GPU - every rank raises after the same collectives, nobody hangs, and the segment is
    removed.  bench.py then keeps the all-gather form of the call."""
    if not os.path.isdir("/dev/shm"):
        pytest.skip("no /dev/shm")
    world, port = 2, 29600 + os.getpid() % 300
    mp.spawn(_shared_vector_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = [open(os.path.join(str(tmp_path), f"r{r}.txt")).read().split() for r in range(world)]
    assert got[0][0] == got[1][0], got
    if not torch.cuda.is_available():
        assert got[0][0] == "OSError"
    assert got[0][1] == got[1][1] == "0", "the shared segment was left behind"
