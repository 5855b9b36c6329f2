"""Row-range sharding logic on CPU: world_size-2 (and 3) gloo process groups.  The per-rank compute is done by the
oracle here (test infrastructure standing in for the GPU kernel); what is under test is the partition, the re-based
per-rank tables, the per-frame broadcast and the gather: concatenated shards == single-table result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from opensubdiv_b200 import shard, synth  # noqa: E402


def test_balanced_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 40, 100_000).astype(np.int32)
    sizes[5000:5100] = 900                                  # a high-valence cluster
    for world in (1, 2, 3, 4, 8):
        ranges = shard.balanced_row_ranges(sizes, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(sizes)
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        plan = shard.ShardPlan(world, 0, ranges)
        assert plan.imbalance(sizes) < 1.02
    assert shard.balanced_row_ranges(np.zeros(0, np.int32), 4) == [(0, 0)] * 4
    ranges = shard.balanced_row_ranges(sizes, 4, align=2048)
    assert all(a % 2048 == 0 for a, _ in ranges[1:])


def test_row_range_tables_are_self_contained():
    mesh = synth.torus_quads(12, 9)
    t = synth.uniform_stencil_table(mesh, 2)
    a, b = 137, 901
    sub = t.row_range(a, b)
    assert sub.num_stencils == b - a and sub.offsets[0] == 0
    assert np.array_equal(np.cumsum(sub.sizes)[:-1], sub.offsets[1:])
    for r in (0, 17, b - a - 1):
        o, n = sub.offsets[r], sub.sizes[r]
        O = t.offsets[a + r]
        assert np.array_equal(sub.indices[o:o + n], t.indices[O:O + n])
        assert np.array_equal(sub.weights[o:o + n], t.weights[O:O + n])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scheme, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    mesh = synth.torus_quads(16, 10) if scheme == "catmark" else synth.torus_tris(14, 9)
    table = synth.uniform_stencil_table(mesh, 2)
    plan = shard.ShardPlan.for_table(table.sizes, world, rank)
    local = shard.local_table(table, plan)
    ncv, L = table.num_control_verts, 3
    bufs = [torch.zeros((ncv, L)), torch.zeros((ncv, L))]
    bc = shard.FrameBroadcaster(bufs, root=0)
    results = []
    for frame in range(3):
        if rank == 0:                                       # only the root knows the deformed control points
            bufs[frame % 2].copy_(torch.from_numpy(synth.deform(mesh.positions, frame)))
        bc.post(frame)
        cv = bc.wait(frame).numpy()
        out = np.zeros((local.num_stencils, L), np.float32)
        assert oracle.eval_stencils(cv.reshape(-1), (0, L, L), [out.reshape(-1)], [(0, L, L)], local.sizes, local.offsets,
                                    local.indices, [local.weights])
        bc.release(frame)
        full = shard.all_gather_rows(torch.from_numpy(out), plan).numpy()
        ref = np.zeros((table.num_stencils, L), np.float32)
        oracle.eval_stencils(synth.deform(mesh.positions, frame).reshape(-1), (0, L, L), [ref.reshape(-1)], [(0, L, L)],
                             table.sizes, table.offsets, table.indices, [table.weights])
        results.append(bool(np.array_equal(full, ref)))
    out_q.put((rank, results, plan.ranges))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,scheme", [(2, "catmark"), (3, "loop")])
def test_sharded_frames_equal_single_table(world, scheme):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scheme, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, results, ranges in got:
        assert results == [True, True, True], (rank, results)
        assert ranges[0][0] == 0


def _worker_locality(rank, world, port, out_q):
    """Strong scaling with the locality plan: every frame the root sends rank r ONLY the control-vertex runs r's rows
    reference (point-to-point, standing in for b200osd_window_pull), every rank applies its own table, the pieces are
    gathered and put back into table order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    mesh = synth.torus_tris(20, 12)
    table = synth.uniform_stencil_table(mesh, 2)
    plans = [shard.LocalityPlan.for_table(table, world, r) for r in range(world)]
    plan = plans[rank]
    local = shard.local_table_rows(table, plan.rows)
    all_runs = [shard.control_runs(shard.local_table_rows(table, p.rows), 16, 4) for p in plans]
    runs = all_runs[rank]
    ncv, L, n = table.num_control_verts, 3, table.num_stencils
    results, received = [], 0
    for frame in range(3):
        cv = torch.full((ncv, L), float("nan"))              # what this rank knows of the frame's control points
        if rank == 0:
            scene = torch.from_numpy(synth.deform(mesh.positions, frame))
            for r in range(1, world):
                for lo, hi in all_runs[r]:
                    dist.send(scene[lo:hi].contiguous(), dst=r)
            for lo, hi in runs:
                cv[lo:hi] = scene[lo:hi]
        else:
            for lo, hi in runs:
                piece = torch.empty((hi - lo, L))
                dist.recv(piece, src=0)
                cv[lo:hi] = piece
        received = sum(hi - lo for lo, hi in runs)
        out = np.zeros((local.num_stencils, L), np.float32)
        assert oracle.eval_stencils(cv.numpy().reshape(-1), (0, L, L), [out.reshape(-1)], [(0, L, L)], local.sizes, local.offsets,
                                    local.indices, [local.weights])
        pieces = [None] * world
        dist.all_gather_object(pieces, (plan.rows, out))
        full = np.zeros((n, L), np.float32)
        for rows, vals in pieces:
            full[rows] = vals
        ref = np.zeros((n, L), np.float32)
        oracle.eval_stencils(synth.deform(mesh.positions, frame).reshape(-1), (0, L, L), [ref.reshape(-1)], [(0, L, L)],
                             table.sizes, table.offsets, table.indices, [table.weights])
        results.append(bool(np.array_equal(full, ref)))
    out_q.put((rank, results, received, ncv))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_locality_sharded_frames_equal_single_table(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_locality, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, results, received, ncv in got:
        assert results == [True, True, True], (rank, results)
        assert received < ncv                               # a rank never needs the whole control mesh
        if world == 4:
            assert received <= 0.6 * ncv


def test_coord_ranges_cover_and_align():
    for n in (0, 1, 31, 32, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            r = shard.coord_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert all(a % 32 == 0 for a, _ in r[1:] if a < n)
            if n >= 32 * world * 8:
                sizes = [b - a for a, b in r]
                assert max(sizes) - min(sizes) <= 64


def _patch_worker(rank, world, port, out_q):
    """EvalPatches sharded by PatchCoord range: replicated tables, per-frame broadcast of the control points, every rank
    refines locally and evaluates its own coordinates; gathered == single evaluation, bit for bit."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from tests.util import golden, table_from, triple_from
    d = golden("patches_catmark_car")
    st, vtx, coords = table_from(d, "st_"), triple_from(d, "vtx_"), d["coords"]
    ncv, nst = st.num_control_verts, st.num_stencils
    plan = shard.coord_plan(len(coords), world, rank)
    mine = np.ascontiguousarray(coords[plan.start:plan.end])
    bufs = [torch.zeros((ncv, 3)), torch.zeros((ncv, 3))]
    bc = shard.FrameBroadcaster(bufs, root=0)
    ok = []

    def evaluate(cv, cs):
        vb = np.zeros((ncv + nst, 3), np.float32)
        vb[:ncv] = cv
        assert oracle.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st.sizes, st.offsets,
                                    st.indices, [st.weights])
        outs = [np.zeros((len(cs), 3), np.float32) for _ in range(3)]
        if len(cs):
            assert oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in outs], [(0, 3, 3)] * 3, cs,
                                       vtx.arrays, vtx.indices, vtx.params)
        return np.concatenate(outs, axis=1)

    for frame in range(2):
        if rank == 0:
            bufs[frame % 2].copy_(torch.from_numpy(synth.deform(d["src0"], frame)))
        bc.post(frame)
        cv = bc.wait(frame).numpy()
        local = evaluate(cv, mine)
        bc.release(frame)
        full = shard.all_gather_rows(torch.from_numpy(local), plan).numpy()
        ok.append(bool(np.array_equal(full, evaluate(synth.deform(d["src0"], frame), coords))))
    out_q.put((rank, ok, plan.ranges))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_patch_coords_equal_single_evaluation():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_patch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ranges in got:
        assert ok == [True, True], (rank, ok)
        assert ranges[0][0] == 0 and ranges[-1][1] > ranges[-1][0]


def test_c_shard_plan_matches_a_straight_restatement():
    """b200osd_shard_plan / b200osd_shard_coords (the C data plane: include/b200osd_capi.h) against a direct numpy
    restatement of their contract; called through ctypes exactly as a C++ host would call them."""
    from opensubdiv_b200 import capi
    rng = np.random.default_rng(7)
    sizes = rng.integers(1, 31, 50_000).astype(np.int32)
    cost = np.cumsum(sizes.astype(np.int64) + 1)
    for world in (1, 2, 5, 8):
        for align in (1, 32, 2048):
            out = np.zeros(2 * world, np.int32)
            assert capi.lib().b200osd_shard_plan(len(sizes), sizes.ctypes.data, world, align, out.ctypes.data) == capi.OK
            cuts = [0]
            for r in range(1, world):
                c = int(np.searchsorted(cost, int(cost[-1]) * r // world, side="left")) + 1
                if align > 1:
                    c = (c + align // 2) // align * align
                cuts.append(min(max(c, cuts[-1]), len(sizes)))
            cuts.append(len(sizes))
            assert out.reshape(world, 2).tolist() == [[cuts[i], cuts[i + 1]] for i in range(world)]
            co = np.zeros(2 * world, np.int64)
            assert capi.lib().b200osd_shard_coords(10_000_001, world, align, co.ctypes.data) == capi.OK
            co = co.reshape(world, 2)
            assert co[0, 0] == 0 and co[-1, 1] == 10_000_001 and (co[1:, 0] == co[:-1, 1]).all()
            assert (np.diff(co, axis=1) >= 0).all() and all(int(a) % align == 0 for a in co[1:, 0])
    bad = np.zeros(2, np.int32)
    assert capi.lib().b200osd_shard_plan(10, None, 1, 1, bad.ctypes.data) == capi.ERR_INVALID


def test_locality_plan_deals_rows_by_the_control_vertices_they_reference():
    """b200osd_shard_plan_locality + b200osd_shard_control_runs (strong scaling of one mesh): the row order is a stable
    sort by the smallest referenced control vertex, the chunks are balanced, every chunk's runs cover what its rows
    reference and are a fraction of the mesh; applying the per-rank tables and scattering the results back equals the
    single-table result (oracle as the per-rank compute)."""
    from oracle import oracle
    mesh = synth.torus_quads(40, 30)
    t = synth.uniform_stencil_table(mesh, 2)
    ncv, n = t.num_control_verts, t.num_stencils
    key = np.minimum.reduceat(t.indices, t.offsets)
    src = np.random.default_rng(3).standard_normal((ncv, 3)).astype(np.float32)
    want = np.zeros((n, 3), np.float32)
    assert oracle.eval_stencils(src.reshape(-1), (0, 3, 3), [want.reshape(-1)], [(0, 3, 3)], t.sizes, t.offsets, t.indices, [t.weights])
    for world in (1, 2, 3, 8):
        plans = [shard.LocalityPlan.for_table(t, world, r) for r in range(world)]
        order = plans[0].row_order
        assert np.array_equal(order, np.argsort(key, kind="stable"))
        assert np.array_equal(np.sort(np.concatenate([p.rows for p in plans])), np.arange(n))
        assert plans[0].imbalance(t.sizes) < 1.02
        got = np.zeros_like(want)
        for p in plans:
            lt = shard.local_table_rows(t, p.rows)
            runs = shard.control_runs(lt, 64, 4)
            assert 1 <= len(runs) <= 4 and all(a < b for a, b in runs)
            covered = np.zeros(ncv, bool)
            for a, b in runs:
                covered[a:b] = True
            assert covered[lt.indices].all()
            assert p.ctrl_lo == lt.indices.min() and p.ctrl_hi == lt.indices.max() + 1
            if world == 8:
                assert covered.sum() < 0.4 * ncv                  # ~1/8 of the mesh + halo (+ the seam's second run)
            # the rank sees only its runs of the control points
            part = np.full_like(src, np.nan)
            part[covered] = src[covered]
            out = np.zeros((len(p.rows), 3), np.float32)
            assert oracle.eval_stencils(part.reshape(-1), (0, 3, 3), [out.reshape(-1)], [(0, 3, 3)], lt.sizes, lt.offsets, lt.indices,
                                        [lt.weights])
            got[p.rows] = out
        assert np.array_equal(got, want)
    from opensubdiv_b200 import capi
    assert capi.lib().b200osd_shard_control_runs(1, None, None, None, 64, 4, np.zeros(8, np.int32).ctypes.data) == -1
    # more pieces than runs allowed: the closest ones merge and still cover
    tbl = synth.SynthStencilTable(100, np.array([1] * 5, np.int32), np.arange(5, dtype=np.int32), np.array([0, 10, 20, 50, 99], np.int32),
                                  np.ones(5, np.float32))
    assert shard.control_runs(tbl, 1, 8) == [(0, 1), (10, 11), (20, 21), (50, 51), (99, 100)]
    assert shard.control_runs(tbl, 1, 3) == [(0, 21), (50, 51), (99, 100)]
    assert shard.control_runs(tbl, 16, 8) == [(0, 32), (48, 64), (96, 100)]


def test_communicator_needs_a_device():
    """No CPU fallback in the exchange either: without a CUDA device b200osd_comm_create returns NULL with a message."""
    import ctypes as C
    from opensubdiv_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    ident = (C.c_char * 128)()
    assert not capi.lib().b200osd_comm_create(1, 0, ident)
    assert capi.last_error()


def test_locality_plan_properties_on_random_tables():
    """Property test (hypothesis) of the host-side partitioning entries on arbitrary small tables: the order is a stable
    sort by the smallest referenced vertex, chunks tile the order, bounding intervals and runs cover every reference, runs
    are sorted, disjoint, at most maxRuns, and never wider than the bounding interval rounded to the granularity."""
    from hypothesis import given, settings, strategies as hs
    from opensubdiv_b200 import capi

    @settings(max_examples=60, deadline=None)
    @given(hs.integers(1, 200), hs.integers(1, 300), hs.integers(1, 6), hs.integers(0, 2 ** 31 - 1), hs.integers(1, 9),
           hs.integers(1, 64), hs.integers(1, 8))
    def check(nrows, ncv, maxsize, seed, world, gran, max_runs):
        rng = np.random.default_rng(seed)
        sizes = rng.integers(0, maxsize + 1, nrows).astype(np.int32)
        offsets = np.zeros(nrows, np.int32)
        offsets[1:] = np.cumsum(sizes[:-1])
        ne = int(sizes.sum())
        indices = rng.integers(0, ncv, max(ne, 1)).astype(np.int32)
        t = synth.SynthStencilTable(ncv, sizes, offsets, indices[:ne], np.ones(ne, np.float32))
        plans = [shard.LocalityPlan.for_table(t, world, r) for r in range(world)]
        key = np.array([indices[offsets[i]:offsets[i] + sizes[i]].min() if sizes[i] else -1 for i in range(nrows)])
        assert np.array_equal(plans[0].row_order, np.argsort(key, kind="stable"))
        assert plans[0].ranges[0][0] == 0 and plans[0].ranges[-1][1] == nrows
        assert all(plans[0].ranges[i][1] == plans[0].ranges[i + 1][0] for i in range(world - 1))
        for p in plans:
            lt = shard.local_table_rows(t, p.rows)
            if lt.num_elements:
                assert p.ctrl_lo == lt.indices.min() and p.ctrl_hi == lt.indices.max() + 1
            runs = shard.control_runs(lt, gran, max_runs)
            assert len(runs) <= max_runs
            assert all(a < b for a, b in runs) and all(runs[i][1] <= runs[i + 1][0] for i in range(len(runs) - 1))
            covered = np.zeros(ncv + gran, bool)
            for a, b in runs:
                covered[a:b] = True
            assert covered[lt.indices].all()
            if lt.num_elements:
                assert runs[0][0] >= (p.ctrl_lo // gran) * gran and runs[-1][1] <= p.ctrl_hi
            else:
                assert runs == []
    check()
