"""-m gpu: the flag-based exchange of the power iteration (csrc/exchange.cu) with two VIRTUAL ranks on one GPU - two
control blocks, two replicas of x, two streams.  The kernels cannot tell peer memory from local memory, so the
protocol (partial sums published to every rank, flags, rank-ordered reduction, pushes of the pieces the other rank
reads, waits) runs exactly as it does over NVLink; arm-spmv_b200/power.py drives the same entry points with one
process per GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class VirtualRank:
    def __init__(self, lib, rank, world, n_total, start, count):
        self.lib, self.rank, self.world, self.start, self.count = lib, rank, world, start, count
        self.ctrl = torch.zeros(lib.thsp_xchg_ctrl_bytes() // 8, dtype=torch.int64, device="cuda")
        self.work = torch.zeros(lib.thsp_xchg_work_bytes() // 4, dtype=torch.int32, device="cuda")
        self.x = torch.full((n_total,), float("nan"), dtype=torch.float64, device="cuda")
        self.ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        self.stream = torch.cuda.Stream()


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("n_total,halo", [(100_003, 517), (4_000_000, 70_000), (64, 64)])
def test_two_virtual_ranks(thsp, cuda, n_total, halo):
    from arm_spmv_b200.lib import check
    lib = thsp.load()
    world = 2
    per = n_total // world
    parts = [(0, per), (per, n_total - per)]
    ranks = [VirtualRank(lib, r, world, n_total, *parts[r]) for r in range(world)]
    ctrl_arr = (C.c_void_p * world)(*[r.ctrl.data_ptr() for r in ranks])
    # rank 0 reads the first `halo` entries of rank 1's slice, rank 1 the last `halo` entries of rank 0's
    h0 = min(halo, parts[1][1])
    h1 = min(halo, parts[0][1])
    dests = {0: (1, per - h1, per), 1: (0, per, per + h0)}
    torch.cuda.synchronize()
    gen = torch.Generator(device="cuda").manual_seed(3)
    for it in range(1, 5):
        ys = [torch.rand(parts[r][1], dtype=torch.float64, device="cuda", generator=gen) - 0.3 for r in range(world)]
        torch.cuda.synchronize()
        for r in ranks:
            with torch.cuda.stream(r.stream):
                s = C.c_void_p(r.stream.cuda_stream)
                check(lib.thsp_xchg_sumsq_publish_f64(C.c_int64(r.count), _p(ys[r.rank]), C.c_uint64(it), world, r.rank, ctrl_arr,
                                                      _p(r.work), s))
        for r in ranks:
            with torch.cuda.stream(r.stream):
                s = C.c_void_p(r.stream.cuda_stream)
                other, lo, hi = dests[r.rank]
                dx = (C.c_void_p * 1)(ranks[other].x.data_ptr())
                dc = (C.c_void_p * 1)(ranks[other].ctrl.data_ptr())
                dlo, dhi = (C.c_int64 * 1)(lo), (C.c_int64 * 1)(hi)
                check(lib.thsp_xchg_scale_push_f64(C.c_int64(r.count), _p(ys[r.rank]), C.c_uint64(it), world, r.rank, _p(r.ctrl),
                                                   _p(r.work), _p(r.x), C.c_int64(r.start), 1, dx, dc, dlo, dhi, _p(r.ss), s))
                check(lib.thsp_xchg_wait(_p(r.ctrl), C.c_uint64(it), C.c_uint(1 << other), s))
        torch.cuda.synchronize()
        tot = sum(float((y.double() ** 2).sum().item()) for y in ys)
        for r in ranks:
            flag = C.c_int(1)
            check(lib.thsp_xchg_timed_out(_p(r.ctrl), C.byref(flag), None))
            assert flag.value == 0
            assert abs(float(r.ss.item()) - tot) <= 1e-13 * tot
        assert ranks[0].ss.item() == ranks[1].ss.item()          # every rank normalises with the same bits
        inv = 1.0 / np.sqrt(ranks[0].ss.item())
        for r in ranks:
            own = r.x[r.start:r.start + r.count]
            assert torch.equal(own, ys[r.rank] * inv)             # vec_axpby's beta == 0 branch: one multiply
            other, lo, hi = dests[r.rank]
            assert torch.equal(ranks[other].x[lo:hi], r.x[lo:hi])  # the pushed piece, bit for bit
            # nothing outside [own slice + pushed piece] of the other replica was touched by this rank
        if it == 1:
            untouched = ranks[1].x[: per - h1]
            assert bool(torch.isnan(untouched).all()) or untouched.numel() == 0


@pytest.mark.parametrize("n,rank,world", [(160, 0, 1), (144, 1, 3)])
def test_host_vector_shared_by_the_ranks(thsp, cuda, oracle, n, rank, world):
    """The distributed y = A x with x in ONE page-locked host vector (power.SharedHostVector, e2e_spmv_step_shared): a
    rank's row blocks pull their windows of x from it through thsp_csr_plan_spmv_host_f64.  One process plays rank
    `rank` of `world`: its slice of y must equal the oracle's rows bit for bit, and only its window of x may have
    been uploaded."""
    from arm_spmv_b200 import power
    ops = power.CudaOps(torch.device("cuda", 0))
    A = power.PartitionedCSR.stencil27(n, rank, world, ops, max_block_rows=3_000_000)
    N = n ** 3
    x = oracle.gen_vector(N, 9)
    xs = power.SharedHostVector(N, 0, 1, torch.device("cuda", 0), tag=f"test{n}")
    try:
        xs.tensor.copy_(torch.from_numpy(x))
        yh = torch.full((A.count,), float("nan"), dtype=torch.float64).pin_memory()

        class It:   # the fields e2e_spmv_step_shared reads
            pass
        it = It()
        it.A, it.ops = A, ops
        it.x = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
        it.y = torch.empty(A.count, dtype=torch.float64, device="cuda")
        for _ in range(3):   # eager, captured, replayed
            yh.fill_(float("nan"))
            power.e2e_spmv_step_shared(it, xs.tensor, yh)
            rp, ci, va = oracle.gen_stencil27_csr(n, A.start, A.start + A.count)
            ref = oracle.csr_spmv(A.count, N, rp, ci, va, x, np.zeros(A.count))
            assert yh.numpy().tobytes() == ref.tobytes()
        up = torch.nonzero(~torch.isnan(it.x.cpu())).flatten()
        lo = min(b.col_min for b in A.blocks); hi = max(b.col_max for b in A.blocks)   # conservative bounds of the blocks
        first, last = int(up[0]), int(up[-1])
        assert up.numel() == last - first + 1, "the uploaded window has holes"
        assert lo <= first <= max(0, A.start - n * n) and min(N - 1, A.start + A.count - 1 + n * n) <= last <= hi
    finally:
        xs.close()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 4096 * 32 - 5, 4096 * 32, 4096 * 32 + 1, 3_000_001])
def test_canonical_sum_of_squares_matches_the_oracle_bits(thsp, cuda, oracle, n):
    """thsp_tile_sumsq_f64 + thsp_tree_sum_f64 (csrc/tree_sum.cuh: 32-row butterfly, index-bit tree) against the
    restatement in oracle.c, bit for bit, around the tile and 4096-tile block sizes; thsp_hash_f64 against its twin,
    and additive over pieces."""
    from arm_spmv_b200.lib import check, current_stream
    lib = thsp.load()
    y = oracle.gen_vector(n, 21) - 0.4
    yd = torch.from_numpy(y).cuda()
    tiles = torch.empty((n + 31) // 32, dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    check(lib.thsp_tile_sumsq_f64(C.c_int64(n), _p(yd), _p(tiles), current_stream()))
    check(lib.thsp_tree_sum_f64(C.c_int64(tiles.numel()), _p(tiles), _p(out), current_stream()))
    want_tiles = oracle.tile_sumsq(y)
    assert tiles.cpu().numpy().tobytes() == want_tiles.tobytes()
    assert float(out.item()) == oracle.tree_sum(want_tiles)
    assert abs(float(out.item()) - float(np.dot(y, y))) <= 1e-13 * float(np.dot(y, y))
    h = C.c_uint64(0)
    check(lib.thsp_hash_f64(C.c_int64(n), _p(yd), C.c_uint64(77), C.byref(h), current_stream()))
    assert h.value == oracle.hash_f64(y, 77)
    cut = n // 3
    h1, h2 = C.c_uint64(0), C.c_uint64(0)
    check(lib.thsp_hash_f64(C.c_int64(cut), _p(yd), C.c_uint64(77), C.byref(h1), current_stream()))
    check(lib.thsp_hash_f64(C.c_int64(n - cut), C.c_void_p(yd.data_ptr() + 8 * cut), C.c_uint64(77 + cut), C.byref(h2), current_stream()))
    assert (h1.value + h2.value) % (1 << 64) == h.value


@pytest.mark.parametrize("kernel", ["stream", "scalar", "merge"])
def test_spmv_epilogue_leaves_the_tile_sums(thsp, cuda, oracle, kernel):
    """thsp_csr_plan_spmv_sumsq_f64: y = A x and, from the same pass (stream kernel) or a pass over y (the others), the
    sum of squares of every 32-row tile - the oracle's bits for the in-order kernels."""
    from arm_spmv_b200 import host as H
    from arm_spmv_b200.lib import check, current_stream
    lib = thsp.load()
    n = 21
    N = n ** 3          # 9261 rows: the last tile is ragged
    A = H.stencil27_csr(n)
    check(lib.thsp_csr_plan_set_kernel(A.plan(), {"scalar": 1, "stream": 3, "merge": 4}[kernel], 1))
    x = H.gen_vector(N, 3)
    y = torch.empty(N, dtype=torch.float64, device="cuda")
    tiles = torch.full(((N + 31) // 32,), float("nan"), dtype=torch.float64, device="cuda")
    check(lib.thsp_csr_plan_spmv_sumsq_f64(A.plan(), _p(x.values), _p(y), 0, _p(tiles), current_stream()))
    torch.cuda.synchronize()
    rp, ci, va = oracle.gen_stencil27_csr(n)
    want = oracle.csr_spmv(N, N, rp, ci, va, oracle.gen_vector(N, 3), np.zeros(N))
    if kernel != "merge":
        assert y.cpu().numpy().tobytes() == want.tobytes()
    assert tiles.cpu().numpy().tobytes() == oracle.tile_sumsq(y.cpu().numpy()).tobytes()


@pytest.mark.parametrize("n_total,halo", [(64 * 4096 * 2, 70_000), (100_032, 517), (64, 32)])
def test_two_virtual_ranks_one_kernel(thsp, cuda, oracle, n_total, halo):
    """thsp_xchg_norm_scale_push_f64 with two virtual ranks on one GPU: the sum is the oracle's canonical sum of the WHOLE
    vector bit for bit (aligned halves are subtrees of one tree), both ranks hold the same bits, the pieces arrive."""
    from arm_spmv_b200.lib import check
    lib = thsp.load()
    world = 2
    per = n_total // world
    parts = [(0, per), (per, n_total - per)]
    ranks = [VirtualRank(lib, r, world, n_total, *parts[r]) for r in range(world)]
    ctrl_arr = (C.c_void_p * world)(*[r.ctrl.data_ptr() for r in ranks])
    h = min(halo, per)
    dests = {0: (1, per - h, per), 1: (0, per, per + h)}
    torch.cuda.synchronize()
    for it in range(1, 4):
        yall = oracle.gen_vector(n_total, 30 + it) - 0.3
        ys = [torch.from_numpy(yall[s:s + c]).cuda() for s, c in parts]
        tiles = [torch.from_numpy(oracle.tile_sumsq(yall[s:s + c])).cuda() for s, c in parts]
        torch.cuda.synchronize()
        for r in ranks:
            with torch.cuda.stream(r.stream):
                s = C.c_void_p(r.stream.cuda_stream)
                other, lo, hi = dests[r.rank]
                dx = (C.c_void_p * 1)(ranks[other].x.data_ptr())
                dc = (C.c_void_p * 1)(ranks[other].ctrl.data_ptr())
                dlo, dhi = (C.c_int64 * 1)(lo), (C.c_int64 * 1)(hi)
                check(lib.thsp_xchg_norm_scale_push_f64(C.c_int64(r.count), _p(ys[r.rank]), _p(tiles[r.rank]), C.c_uint64(it), world, r.rank,
                                                        ctrl_arr, _p(r.work), _p(r.x), C.c_int64(r.start), 1, dx, dc, dlo, dhi, _p(r.ss), s))
                check(lib.thsp_xchg_wait(_p(r.ctrl), C.c_uint64(it), C.c_uint(1 << other), s))
        torch.cuda.synchronize()
        want = oracle.tree_sum(oracle.tile_sumsq(yall))
        for r in ranks:
            flag = C.c_int(1)
            check(lib.thsp_xchg_timed_out(_p(r.ctrl), C.byref(flag), None))
            assert flag.value == 0
            if per % 32 == 0 and (per // 32) & (per // 32 - 1) == 0:
                assert float(r.ss.item()) == want     # the halves are subtrees of the whole vector's tree
            assert abs(float(r.ss.item()) - want) <= 1e-13 * want
        assert ranks[0].ss.item() == ranks[1].ss.item()
        inv = 1.0 / np.sqrt(ranks[0].ss.item())
        for r in ranks:
            own = r.x[r.start:r.start + r.count]
            assert torch.equal(own, ys[r.rank] * inv)
            other, lo, hi = dests[r.rank]
            assert torch.equal(ranks[other].x[lo:hi], r.x[lo:hi])


# ---- normalisation deferred into the next product (power.DeferredPowerIteration) -------------------------------------
@pytest.mark.parametrize("n,scale", [(21, 0.3721), (40, 1.0), (33, 7.0e-3)])
def test_scaled_spmv_is_the_spmv_of_the_scaled_vector(thsp, cuda, oracle, n, scale):
    """thsp_csr_plan_spmv_scaled_f64: y = A (s x) with s applied to every gathered x_j inside the stream kernel has the bits
    of the product with a vector that was scaled by a pass first (thsp_scale_by_dev_f64) - and of the oracle's serial
    product with s * x; the tile sums of squares come out of the same epilogue."""
    from arm_spmv_b200 import host as H
    from arm_spmv_b200.lib import check, current_stream
    lib = thsp.load()
    N = n ** 3
    A = H.stencil27_csr(n)
    check(lib.thsp_csr_plan_set_kernel(A.plan(), 3, 1))
    x = H.gen_vector(N, 9)
    s = torch.tensor([scale], dtype=torch.float64, device="cuda")
    y1 = torch.empty(N, dtype=torch.float64, device="cuda")
    y2 = torch.empty(N, dtype=torch.float64, device="cuda")
    t1 = torch.full(((N + 31) // 32,), float("nan"), dtype=torch.float64, device="cuda")
    xs = torch.empty(N, dtype=torch.float64, device="cuda")
    check(lib.thsp_csr_plan_spmv_scaled_f64(A.plan(), _p(x.values), _p(s), _p(y1), 0, _p(t1), current_stream()))
    check(lib.thsp_scale_by_dev_f64(C.c_int64(N), _p(x.values), _p(s), _p(xs), current_stream()))
    check(lib.thsp_csr_plan_spmv_f64(A.plan(), _p(xs), _p(y2), 0, current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    rp, ci, va = oracle.gen_stencil27_csr(n)
    xh = oracle.gen_vector(N, 9)
    assert xs.cpu().numpy().tobytes() == (xh * scale).tobytes()
    want = oracle.csr_spmv(N, N, rp, ci, va, xh * scale, np.zeros(N))
    assert y1.cpu().numpy().tobytes() == want.tobytes()
    assert t1.cpu().numpy().tobytes() == oracle.tile_sumsq(want).tobytes()
    # a plan that runs another kernel refuses instead of silently ignoring the factor
    check(lib.thsp_csr_plan_set_kernel(A.plan(), 1, 1))
    assert lib.thsp_csr_plan_spmv_scaled_f64(A.plan(), _p(x.values), _p(s), _p(y1), 0, _p(t1), current_stream()) != 0


@pytest.mark.parametrize("n_total,halo", [(64 * 4096 * 2, 70_001), (100_032, 517), (64, 32)])
def test_two_virtual_ranks_deferred(thsp, cuda, oracle, n_total, halo):
    """thsp_xchg_norm_push_f64 with two virtual ranks on one GPU: the canonical sum of the whole vector and its 1/sqrt on
    both ranks (same bits), the RAW pieces in the other rank's vector, everything else of that vector untouched."""
    from arm_spmv_b200.lib import check
    lib = thsp.load()
    world = 2
    per = n_total // world
    parts = [(0, per), (per, n_total - per)]
    ranks = [VirtualRank(lib, r, world, n_total, *parts[r]) for r in range(world)]
    invs = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
    ctrl_arr = (C.c_void_p * world)(*[r.ctrl.data_ptr() for r in ranks])
    h = min(halo, per)
    dests = {0: (1, per - h, per), 1: (0, per, per + h)}
    torch.cuda.synchronize()
    for it in range(1, 4):
        yall = oracle.gen_vector(n_total, 40 + it) - 0.3
        tiles = [torch.from_numpy(oracle.tile_sumsq(yall[s:s + c])).cuda() for s, c in parts]
        for r in ranks:   # a rank's rows of y sit in its own slice of its vector
            r.x.fill_(float("nan"))
            r.x[r.start:r.start + r.count] = torch.from_numpy(yall[r.start:r.start + r.count]).cuda()
        torch.cuda.synchronize()
        for r in ranks:
            with torch.cuda.stream(r.stream):
                s = C.c_void_p(r.stream.cuda_stream)
                other, lo, hi = dests[r.rank]
                dx = (C.c_void_p * 1)(ranks[other].x.data_ptr())
                dc = (C.c_void_p * 1)(ranks[other].ctrl.data_ptr())
                dlo, dhi = (C.c_int64 * 1)(lo), (C.c_int64 * 1)(hi)
                yown = C.c_void_p(r.x.data_ptr() + 8 * r.start)
                check(lib.thsp_xchg_norm_push_f64(C.c_int64(r.count), yown, _p(tiles[r.rank]), C.c_uint64(it), world, r.rank, ctrl_arr,
                                                  _p(r.work), C.c_int64(r.start), 1, dx, dc, dlo, dhi, _p(r.ss), _p(invs[r.rank]), s))
                check(lib.thsp_xchg_wait(_p(r.ctrl), C.c_uint64(it), C.c_uint(1 << other), s))
        torch.cuda.synchronize()
        want = oracle.tree_sum(oracle.tile_sumsq(yall))
        for r in ranks:
            flag = C.c_int(1)
            check(lib.thsp_xchg_timed_out(_p(r.ctrl), C.byref(flag), None))
            assert flag.value == 0
            if per % 32 == 0 and (per // 32) & (per // 32 - 1) == 0:
                assert float(r.ss.item()) == want
            assert abs(float(r.ss.item()) - want) <= 1e-13 * want
            assert float(invs[r.rank].item()) == 1.0 / np.sqrt(float(r.ss.item()))
        assert ranks[0].ss.item() == ranks[1].ss.item()
        for r in ranks:
            got = r.x.cpu().numpy()
            other, lo, hi = dests[1 - r.rank][0], dests[1 - r.rank][1], dests[1 - r.rank][2]   # what the OTHER rank pushed here
            exp = np.full(n_total, np.nan)
            exp[r.start:r.start + r.count] = yall[r.start:r.start + r.count]
            exp[lo:hi] = yall[lo:hi]
            assert got.tobytes() == exp.tobytes()
    # a piece outside the own slice is refused
    r = ranks[0]
    dx = (C.c_void_p * 1)(ranks[1].x.data_ptr())
    dc = (C.c_void_p * 1)(ranks[1].ctrl.data_ptr())
    bad_lo, bad_hi = (C.c_int64 * 1)(per - 1), (C.c_int64 * 1)(per + 1)
    assert lib.thsp_xchg_norm_push_f64(C.c_int64(r.count), C.c_void_p(r.x.data_ptr()), _p(tiles[0]), C.c_uint64(9), world, 0, ctrl_arr, _p(r.work),
                                       C.c_int64(0), 1, dx, dc, bad_lo, bad_hi, _p(r.ss), _p(invs[0]), None) != 0


@pytest.mark.parametrize("n,generic", [(32, False), (24, False), (12, True)])
def test_deferred_loop_has_the_bits_of_the_eager_loop(thsp, cuda, oracle, n, generic):
    """One GPU: DeferredPowerIteration (no normalising pass, the factor folded into the next SpMV) against PowerIteration
    (SpMV, sum, x = y / ||y||) - norm, y and the materialised x equal bit for bit after every step, and equal to the
    oracle's serial loop with the canonical sum."""
    from arm_spmv_b200 import power
    ops = power.CudaOps(cuda)
    if generic:   # short matrix through from_csr: the plan picks a non-stream kernel -> scaled copy of x
        rp, ci, va = oracle.gen_stencil27_csr(n)
        mk = lambda: power.PartitionedCSR.from_csr(n ** 3, torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda(), torch.from_numpy(va).cuda(), 0, 1, ops)
    else:
        mk = lambda: power.PartitionedCSR.stencil27(n, 0, 1, ops)
    Aa, Ab = mk(), mk()
    if generic:
        from arm_spmv_b200.lib import check
        for blk in Aa.blocks + Ab.blocks:
            check(thsp.load().thsp_csr_plan_set_kernel(blk.payload.plan(), 1, 1))
            assert blk.payload.plan_kernel()[0] == "scalar"
    a = power.make_iteration(Aa, ops, exchange="allgather", seed=5)
    b = power.make_iteration(Ab, ops, exchange="deferred", seed=5)
    assert isinstance(b, power.DeferredPowerIteration)
    N = n ** 3
    rp, ci, va = oracle.gen_stencil27_csr(n)
    xs = oracle.gen_vector(N, 5)
    for _ in range(4):
        a.step()
        b.step()
        torch.cuda.synchronize()
        ys = oracle.csr_spmv(N, N, rp, ci, va, xs, np.zeros(N))
        nrm = np.sqrt(oracle.tree_sum(oracle.tile_sumsq(ys)))
        xs = oracle.axpby(1.0 / nrm, ys, 0.0, ys)
        assert a.norm() == b.norm() == nrm
        assert torch.equal(a.y, b.y) and a.y_hash() == b.y_hash()
        assert a.y.cpu().numpy().tobytes() == ys.tobytes()
        assert torch.equal(a.x, b.x)
        assert b.x.cpu().numpy().tobytes() == xs.tobytes()
