"""EvalStencils parity on a real B200: CUDA path (through the C ABI) vs the oracle and the reference's golden outputs."""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from opensubdiv_b200 import synth
from tests.gpu_util import D, dev, oracle_stencils
from tests.util import golden, golden_names, table_from, assert_close, REL_TOL

pytestmark = pytest.mark.gpu
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")
# 0 auto, 1 CSR kernel on the table, 2 scalar gathers, 8 persistent grid, 11 64 warps/SM, 12 both,
# 1CS: streams staged by TMA bulk copies into per-warp shared-memory rings, C groups per chunk, S stages
VARIANTS = (0, 1, 2, 8, 11, 12, 113, 122, 143)


def refine_same_buffer(t, src, L, variant=0, idx16=True, sort_elements=False):
    """Osd::Mesh::Refine layout: one buffer [control | refined], src and dst descriptors into it (osd/mesh.h:505-519)."""
    ncv, n = t.num_control_verts, t.num_stencils
    vb = osd.B200VertexBuffer.Create(L, ncv + n)
    vb.UpdateData(np.ascontiguousarray(src, np.float32), 0, ncv)
    tbl = osd.B200StencilTable.Create(t, idx16=idx16, sort_elements=sort_elements)
    assert tbl is not None and tbl.GetNumStencils() == n
    tbl.SetVariant(variant)
    try:
        assert osd.B200Evaluator.EvalStencils(vb, D(0, L, L), vb, D(ncv * L, L, L), tbl)
    finally:
        tbl.SetVariant(0)
    osd.B200Evaluator.Synchronize()
    return vb.as_tensor()[ncv:].cpu().numpy()


def test_config1_catmark_cube_level4():
    """BASELINE config 1: catmark_cube, uniform level 4, float3 xyz -- vs CpuEvaluator output and the Hbr golden file."""
    d = golden("catmark_cube_L4")
    for which in ("last_", "all_"):
        t = table_from(d, which)
        scale = oracle_stencils(d["src"], (0, 3, 3), t.num_stencils, 3, t, 1, abs_scale=True)[0]
        for v in VARIANTS:
            out = refine_same_buffer(t, d["src"], 3, v)
            assert_close(out, d[which + "out"], scale, f"cube {which} variant {v}")
    out = refine_same_buffer(table_from(d, "last_"), d["src"], 3).astype(np.float64)
    hbr = d["hbr_level3"].astype(np.float64)
    dist = np.sqrt(((out[:, None, :] - hbr[None, :, :]) ** 2).sum(-1)).min(axis=1)
    assert dist.max() <= 1e-6          # the reference's own tolerance (regression/osd_regression/main.cpp:53)


@pytest.mark.parametrize("name", golden_names("stencils_"))
def test_regression_shapes_vertex_and_varying(name):
    d = golden(name)
    L = d["src"].shape[1]
    for prefix, key in (("t_", "out"), ("v_", "v_out")):
        t = table_from(d, prefix)
        scale = oracle_stencils(d["src"], (0, L, L), t.num_stencils, L, t, 1, abs_scale=True)[0]
        for v in VARIANTS:
            assert_close(refine_same_buffer(t, d["src"], L, v), d[key], scale, f"{name} {prefix} variant {v}")
        for v in (0, 8):      # the same table kept with 32-bit indices / with the reference's element order
            assert_close(refine_same_buffer(t, d["src"], L, v, idx16=False), d[key], scale, f"{name} {prefix} idx32 variant {v}")
            # opt-in element sorting changes the summation order: same terms, looser agreement on rows of 100+ terms
            assert_close(refine_same_buffer(t, d["src"], L, v, sort_elements=True), d[key], scale, f"{name} {prefix} sorted variant {v}",
                         tol=5e-6)


@pytest.mark.parametrize("name", golden_names("limit_"))
@pytest.mark.parametrize("nw", [1, 3, 6])
def test_limit_stencils_with_derivatives(name, nw):
    """LimitStencilTable with du/dv/duu/duv/dvv, outputs interleaved in ONE buffer addressed by different descriptors
    (examples/glStencilViewer/glStencilViewer.cpp:188-191)."""
    d = golden(name)
    t = table_from(d, "t_")
    n = t.num_stencils
    src = dev(d["src"])
    tbl = osd.B200StencilTable.Create(t)
    scales = oracle_stencils(d["src"], (0, 3, 3), n, 3, t, nw, abs_scale=True)
    for v in VARIANTS:
        out = torch.full((n, 3 * nw), float("nan"), device="cuda")
        args = []
        for k in range(nw):
            args += [out, D(3 * k, 3, 3 * nw)]
        tbl.SetVariant(v)
        assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), *args, tbl)
        tbl.SetVariant(0)
        res = out.cpu().numpy()
        for k in range(nw):
            assert_close(res[:, 3 * k:3 * k + 3], d["out_" + OUT6[k]], scales[k], f"{name} {OUT6[k]} variant {v}")
    # raw-pointer overload on the reference-layout device arrays (osd/cudaEvaluator.h:449-466)
    outs = [torch.zeros((n, 3), device="cuda") for _ in range(nw)]
    ws = [tbl.GetWeightsBuffer(), tbl.GetDuWeightsBuffer(), tbl.GetDvWeightsBuffer(), tbl.GetDuuWeightsBuffer(),
          tbl.GetDuvWeightsBuffer(), tbl.GetDvvWeightsBuffer()][:nw]
    assert osd.B200Evaluator.EvalStencilsRaw(src, D(0, 3, 3), [(o, D(0, 3, 3)) for o in outs], tbl.GetSizesBuffer(),
                                             tbl.GetOffsetsBuffer(), tbl.GetIndicesBuffer(), ws, 0, n)
    for k in range(nw):
        assert_close(outs[k].cpu().numpy(), d["out_" + OUT6[k]], scales[k], f"{name} raw {OUT6[k]}")
    # a client that only holds those device arrays (an Osd::CudaStencilTable) converts them once and gets the fast path:
    # bit-identical to the table created from the host arrays
    ws9 = [tbl.GetWeightsBuffer(), tbl.GetDuWeightsBuffer(), tbl.GetDvWeightsBuffer(), tbl.GetDuuWeightsBuffer(),
           tbl.GetDuvWeightsBuffer(), tbl.GetDvvWeightsBuffer()]
    ws9 = [w if k < nw else None for k, w in enumerate(ws9)]
    tbl2 = osd.B200StencilTable.CreateFromDevice(n, tbl.GetSizesBuffer(), tbl.GetOffsetsBuffer(), tbl.GetIndicesBuffer(), *ws9,
                                                 numControlVertices=t.num_control_verts)
    assert tbl2.GetNumStencils() == n
    a = torch.full((n, 3 * nw), float("nan"), device="cuda")
    b = torch.full((n, 3 * nw), float("nan"), device="cuda")
    aa, bb = [], []
    for k in range(nw):
        aa += [a, D(3 * k, 3, 3 * nw)]
        bb += [b, D(3 * k, 3, 3 * nw)]
    assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), *aa, tbl)
    assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), *bb, tbl2)
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", golden_names("stencils_") + golden_names("limit_"))
def test_reference_exact_mode_is_bit_identical_to_the_cpu_evaluator(name):
    """flag 64 (reference_exact): a rounded product added to the running sum in the table's order, what osd/cpuKernel.cpp
    does -- the committed outputs of Osd::CpuEvaluator come out bit for bit, for values and all derivative streams, through
    the table and through a row range."""
    d = golden(name)
    t = table_from(d, "t_")
    n = t.num_stencils
    nw = 6 if t.dvv is not None else (3 if t.du is not None else 1)
    src = dev(d["src"])
    L = d["src"].shape[1]
    keys = ["out"] if "out" in d.files else ["out_" + k for k in OUT6]
    tbl = osd.B200StencilTable.Create(t, reference_exact=True)
    for k in (1, 3, 6):
        if k > nw or k > len(keys):
            continue
        outs = [torch.full((n, L), float("nan"), device="cuda") for _ in range(k)]
        args = []
        for o in outs:
            args += [o, D(0, L, L)]
        assert osd.B200Evaluator.EvalStencils(src, D(0, L, L), *args, tbl)
        for q in range(k):
            assert np.array_equal(outs[q].cpu().numpy().view(np.int32), d[keys[q]].view(np.int32)), (name, keys[q], k)
    a, b = n // 3, n - n // 4
    out = torch.zeros((n, L), device="cuda")
    assert osd.B200Evaluator.EvalStencils(src, D(0, L, L), out, D(0, L, L), tbl, start=a, end=b)
    assert np.array_equal(out.cpu().numpy()[a:b].view(np.int32), d[keys[0]][a:b].view(np.int32))


@pytest.mark.parametrize("name", golden_names("stencils_") + golden_names("limit_"))
def test_bucketed_layout_built_on_the_device_equals_the_host_builder(name):
    """The bucketed copy is laid out by kernels from the uploaded arrays (sell_extent_kernel / sell_fill_kernel); flag 32
    keeps the host builder.  Same slices, same bytes: evaluations are bit-identical and stream the same number of bytes,
    with 16- and 32-bit slice indices, in the default and in the table's own summation order, for 1 / 3 / 6 weight streams."""
    from opensubdiv_b200 import capi
    d = golden(name)
    t = table_from(d, "t_")
    n = t.num_stencils
    nw = 6 if t.dvv is not None else (3 if t.du is not None else 1)
    src = dev(d["src"])
    L = d["src"].shape[1]
    for kw in ({}, {"keep_order": True}, {"idx16": False}):
        a_tbl = osd.B200StencilTable.Create(t, **kw)
        b_tbl = osd.B200StencilTable.Create(t, host_layout=True, **kw)
        for k in (1, 3, 6):
            if k > nw:
                continue
            assert capi.lib().b200osd_stencil_table_stream_bytes(a_tbl._h, k) == capi.lib().b200osd_stencil_table_stream_bytes(b_tbl._h, k)
            a = torch.full((n, L * k), float("nan"), device="cuda")
            b = torch.full((n, L * k), float("nan"), device="cuda")
            aa, bb = [], []
            for q in range(k):
                aa += [a, D(L * q, L, L * k)]
                bb += [b, D(L * q, L, L * k)]
            assert osd.B200Evaluator.EvalStencils(src, D(0, L, L), *aa, a_tbl)
            assert osd.B200Evaluator.EvalStencils(src, D(0, L, L), *bb, b_tbl)
            assert torch.equal(a, b), (name, kw, k)


@pytest.mark.parametrize("L,stride,offset", [(1, 1, 0), (2, 2, 0), (3, 3, 0), (4, 4, 0), (5, 5, 0), (6, 6, 0), (8, 8, 0),
                                             (12, 12, 0), (3, 4, 1), (6, 9, 2), (4, 8, 4), (3, 7, 2), (1, 5, 4), (17, 20, 1)])
def test_descriptors_lengths_strides_offsets(L, stride, offset):
    d = golden("stencils_catmark_car")
    t = table_from(d, "t_")
    ncv, n = t.num_control_verts, t.num_stencils
    rng = np.random.default_rng(L * 100 + stride)
    src = rng.standard_normal(offset + ncv * stride + 4).astype(np.float32)
    expect = np.full(offset + n * stride + 4, np.nan, np.float32)
    from oracle import oracle
    assert oracle.eval_stencils(src, (offset, L, stride), [expect], [(offset, L, stride)], t.sizes, t.offsets, t.indices,
                                [t.weights])
    scale = np.full_like(expect, np.nan)
    with oracle.abs_mode():
        oracle.eval_stencils(src, (offset, L, stride), [scale], [(offset, L, stride)], t.sizes, t.offsets, t.indices, [t.weights])
    tbl = osd.B200StencilTable.Create(t)
    for v in VARIANTS:
        out = torch.full((len(expect),), float("nan"), device="cuda")
        tbl.SetVariant(v)
        assert osd.B200Evaluator.EvalStencils(dev(src), D(offset, L, stride), out, D(offset, L, stride), tbl)
        tbl.SetVariant(0)
        got = out.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(expect)), "wrote outside the described elements"
        m = ~np.isnan(expect)
        assert_close(got[m], expect[m], scale[m], f"L={L} stride={stride} off={offset} variant {v}")


def test_row_ranges_noop_and_errors():
    d = golden("stencils_catmark_car")
    t = table_from(d, "t_")
    n = t.num_stencils
    src = dev(d["src"])
    tbl = osd.B200StencilTable.Create(t)
    full = oracle_stencils(d["src"], (0, 3, 3), n, 3, t, 1)[0]
    scale = oracle_stencils(d["src"], (0, 3, 3), n, 3, t, 1, abs_scale=True)[0]
    for (a, b) in [(0, n), (0, 1), (n - 1, n), (17, 4099), (2048, 4096), (2047, 2049), (5000, 5001)]:
        for v in VARIANTS:
            out = torch.full((n, 3), float("nan"), device="cuda")
            tbl.SetVariant(v)
            assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), out, D(0, 3, 3), tbl, start=a, end=b)
            tbl.SetVariant(0)
            got = out.cpu().numpy()
            assert np.isnan(got[:a]).all() and np.isnan(got[b:]).all(), "rows outside [start,end) were written"
            assert_close(got[a:b], full[a:b], scale[a:b], f"range {a}:{b} variant {v}")     # absolute row addressing
    out = torch.full((n, 3), float("nan"), device="cuda")
    assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), out, D(0, 3, 3), tbl, start=9, end=9)      # no-op -> true
    assert torch.isnan(out).all()
    assert not osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), out, D(0, 4, 4), tbl)                  # length mismatch -> false
    assert not osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), None, D(0, 3, 3), tbl)                 # NULL dst -> false
    assert torch.isnan(out).all()


def test_empty_and_degenerate_tables():
    empty = type("T", (), dict(sizes=np.zeros(0, np.int32), offsets=np.zeros(0, np.int32), indices=np.zeros(0, np.int32),
                               weights=np.zeros(0, np.float32)))()
    tbl = osd.B200StencilTable.Create(empty)
    assert tbl is not None and tbl.GetNumStencils() == 0
    out = torch.zeros(8, device="cuda")
    assert osd.B200Evaluator.EvalStencils(out, D(0, 3, 3), out, D(0, 3, 3), tbl)
    # rows of size 0 (weights sum to nothing -> zeros) mixed with a row of 300 terms
    rng = np.random.default_rng(0)
    sizes = np.array([0, 300, 0, 1, 33, 0], np.int32)
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    ne = int(sizes.sum())
    t = type("T", (), dict(num_control_verts=50, num_stencils=6, sizes=sizes, offsets=offsets,
                           indices=rng.integers(0, 50, ne).astype(np.int32),
                           weights=rng.standard_normal(ne).astype(np.float32), du=None, dv=None, duu=None, duv=None, dvv=None))()
    src = rng.standard_normal((50, 3)).astype(np.float32)
    exp = oracle_stencils(src, (0, 3, 3), 6, 3, t, 1)[0]
    scale = oracle_stencils(src, (0, 3, 3), 6, 3, t, 1, abs_scale=True)[0]
    tbl = osd.B200StencilTable.Create(t)
    for v in VARIANTS:
        out = torch.full((6, 3), float("nan"), device="cuda")
        tbl.SetVariant(v)
        assert osd.B200Evaluator.EvalStencils(dev(src), D(0, 3, 3), out, D(0, 3, 3), tbl)
        tbl.SetVariant(0)
        assert_close(out.cpu().numpy(), exp, np.maximum(scale, 1e-6), f"degenerate variant {v}")


def test_wide_index_span_falls_back_to_32bit_indices():
    """Slices whose rows reference control vertices > 65535 apart cannot use 16-bit offsets: the table must silently keep
    32-bit indices (and give the same results)."""
    rng = np.random.default_rng(3)
    ncv, n = 300_000, 5000
    sizes = rng.integers(1, 24, n).astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    ne = int(sizes.sum())
    t = type("T", (), dict(num_control_verts=ncv, num_stencils=n, sizes=sizes, offsets=offsets,
                           indices=rng.integers(0, ncv, ne).astype(np.int32), weights=rng.random(ne).astype(np.float32),
                           du=None, dv=None, duu=None, duv=None, dvv=None))()
    src = rng.standard_normal((ncv, 3)).astype(np.float32)
    exp = oracle_stencils(src, (0, 3, 3), n, 3, t, 1)[0]
    scale = oracle_stencils(src, (0, 3, 3), n, 3, t, 1, abs_scale=True)[0]
    tbl = osd.B200StencilTable.Create(t)
    assert tbl.GetStreamBytes(1) > int(sizes.sum()) * 8          # 4-byte indices + 4-byte weights (+ padding)
    for v in VARIANTS:
        out = torch.full((n, 3), float("nan"), device="cuda")
        tbl.SetVariant(v)
        assert osd.B200Evaluator.EvalStencils(dev(src), D(0, 3, 3), out, D(0, 3, 3), tbl)
        tbl.SetVariant(0)
        assert_close(out.cpu().numpy(), exp, scale, f"wide span variant {v}")


@pytest.mark.parametrize("L,num_instances", [(3, 1), (3, 2), (3, 7), (6, 5), (4, 4), (5, 3)])
def test_batched_instances_equal_per_instance_calls(L, num_instances):
    """One topology, many control-point sets in ONE buffer at a constant vertex pitch (examples/glShareTopology layout:
    each instance = [its control vertices | its refined vertices]); the batched call must reproduce the reference's
    one-EvalStencils-per-instance loop bit for bit."""
    d = golden("stencils_catmark_car")
    t = table_from(d, "t_")
    ncv, n = t.num_control_verts, t.num_stencils
    pitch = (ncv + n) * L                                    # floats between instances
    rng = np.random.default_rng(L * 10 + num_instances)
    host = np.zeros((num_instances, ncv + n, L), np.float32)
    host[:, :ncv] = rng.standard_normal((num_instances, ncv, L)).astype(np.float32)
    tbl = osd.B200StencilTable.Create(t)
    a = dev(host.reshape(-1))
    b = dev(host.reshape(-1))
    for inst in range(num_instances):                        # reference pattern: shifted descriptors, one call each
        assert osd.B200Evaluator.EvalStencils(a, D(inst * pitch, L, L), a, D(inst * pitch + ncv * L, L, L), tbl)
    assert osd.B200Evaluator.EvalStencilsBatched(b, D(0, L, L), b, D(ncv * L, L, L), tbl, num_instances, pitch)
    assert torch.equal(a, b)
    exp = oracle_stencils(host[-1, :ncv], (0, L, L), n, L, t, 1)[0]
    scale = oracle_stencils(host[-1, :ncv], (0, L, L), n, L, t, 1, abs_scale=True)[0]
    got = b.view(num_instances, ncv + n, L)[-1, ncv:].cpu().numpy()
    assert_close(got, exp, scale, f"batched L={L} x{num_instances}")
    # row sub-range + errors behave like the single-instance call
    c = dev(host.reshape(-1))
    assert osd.B200Evaluator.EvalStencilsBatched(c, D(0, L, L), c, D(ncv * L, L, L), tbl, num_instances, pitch, start=100, end=3000)
    cv = c.view(num_instances, ncv + n, L)
    assert torch.equal(cv[:, ncv + 100:ncv + 3000], b.view(num_instances, ncv + n, L)[:, ncv + 100:ncv + 3000])
    assert (cv[:, ncv:ncv + 100] == 0).all() and (cv[:, ncv + 3000:] == 0).all()
    assert not osd.B200Evaluator.EvalStencilsBatched(c, D(0, L, L), c, D(0, L + 1, L + 1), tbl, num_instances, pitch)


@pytest.fixture(scope="module")
def config2():
    """BASELINE config 2 at full size: Catmark torus 400x250 (100k control verts), uniform level 3, last level:
    6.4 M rows / 84.1 M elements, 6-float interleaved xyz+normal."""
    mesh = synth.torus_quads(400, 250)
    table = synth.uniform_stencil_table(mesh, 3)
    assert table.num_stencils == 6_400_000 and table.num_elements == 84_100_000
    tbl = osd.B200StencilTable.Create(table)
    assert tbl is not None
    return mesh, table, tbl


def _prim6(mesh, frame):
    p = synth.deform(mesh.positions, frame)
    return np.ascontiguousarray(np.concatenate([p, synth.vertex_normals_like(p)], axis=1), np.float32)


@pytest.mark.slow
def test_config2_full_size_frames_vs_oracle(config2):
    """Every one of the 6.4 M rows of three frames against the reference's own Osd::CpuEvaluator (oracle/_ref, when it is
    on the box; the C oracle otherwise), relative to the scale sum|w||x| of every row."""
    from oracle import oracle, ref
    mesh, table, tbl = config2
    ncv, n = table.num_control_verts, table.num_stencils
    vb = osd.B200VertexBuffer.Create(6, ncv + n)
    worst = 0.0
    for frame in (0, 1, 17):
        src = _prim6(mesh, frame)
        vb.UpdateData(src, 0, ncv)
        assert osd.B200Evaluator.EvalStencils(vb, D(0, 6, 6), vb, D(ncv * 6, 6, 6), tbl)
        osd.B200Evaluator.Synchronize()
        got = vb.as_tensor()[ncv:].cpu().numpy()
        exp = np.zeros((n, 6), np.float32)
        scl = np.zeros((n, 6), np.float32)
        if ref.available():
            assert ref.eval_stencils(src.reshape(-1), (0, 6, 6), [exp.reshape(-1)], [(0, 6, 6)], table, impl="cpu")
        else:
            assert oracle.eval_stencils(src.reshape(-1), (0, 6, 6), [exp.reshape(-1)], [(0, 6, 6)], table.sizes, table.offsets,
                                        table.indices, [table.weights])
        with oracle.abs_mode():
            oracle.eval_stencils(src.reshape(-1), (0, 6, 6), [scl.reshape(-1)], [(0, 6, 6)], table.sizes, table.offsets,
                                 table.indices, [table.weights])
        worst = max(worst, assert_close(got, exp, scl, f"frame {frame}, all {n} rows"))
    print(f"CONFIG2 all rows x 3 frames vs Osd::CpuEvaluator: worst relative error {worst:.3e}")


@pytest.mark.slow
def test_config2_table_built_by_far_reference_order_vs_sorted():
    """Config 2 with the table built by the REAL Far::StencilTableFactory (insertion order inside rows, unlike the
    index-sorted synthetic table): parity against the oracle in the reference's order, and the effect of the opt-in
    element sorting on speed (printed; recorded in profiles/)."""
    from oracle import oracle, ref
    if not ref.available():
        pytest.skip("oracle/_ref/libosdref.so not present")
    mesh = synth.torus_quads(400, 250)
    m = ref.Mesh.from_topology("catmark", mesh.num_verts, np.full(len(mesh.faces), 4, np.int32), mesh.faces.reshape(-1))
    far = m.refine_uniform(3).stencil_table()
    assert far.num_stencils == 6_400_000
    ncv, n = far.num_control_verts, far.num_stencils
    src = _prim6(mesh, 2)
    x = dev(src)
    out = torch.empty((n, 6), device="cuda")
    import time
    for sort, kw in (("reference order", dict(keep_order=True)), ("rows<=16 sorted (default)", {}), ("all rows sorted", dict(sort_elements=True))):
        tbl = osd.B200StencilTable.Create(far, **kw)
        for _ in range(5):
            assert osd.B200Evaluator.EvalStencils(x, D(0, 6, 6), out, D(0, 6, 6), tbl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            osd.B200Evaluator.EvalStencils(x, D(0, 6, 6), out, D(0, 6, 6), tbl)
        e1.record()
        torch.cuda.synchronize()
        print(f"FAR-ORDER-TABLE cfg2 L=6 {sort}: {e0.elapsed_time(e1) / 50:.4f} ms/frame, stream {tbl.GetStreamBytes(1) / 1e6:.0f} MB")
        # all 6.4 M rows against Osd::CpuEvaluator on the same Far table
        exp = np.zeros((n, 6), np.float32)
        scl = np.zeros((n, 6), np.float32)
        assert ref.eval_stencils(src.reshape(-1), (0, 6, 6), [exp.reshape(-1)], [(0, 6, 6)], far, impl="cpu")
        with oracle.abs_mode():
            oracle.eval_stencils(src.reshape(-1), (0, 6, 6), [scl.reshape(-1)], [(0, 6, 6)], far.sizes, far.offsets, far.indices,
                                 [far.weights])
        worst = assert_close(out.cpu().numpy(), exp, scl, f"far table {sort}")
        print(f"FAR-ORDER-TABLE cfg2 L=6 {sort}: worst relative error over all rows {worst:.3e}")
        del tbl


@pytest.mark.slow
def test_config2_size_independent_properties(config2):
    mesh, table, tbl = config2
    ncv, n = table.num_control_verts, table.num_stencils
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((ncv, 6), device="cuda", generator=g)
    y = torch.randn((ncv, 6), device="cuda", generator=g)

    def ev(src, variant=0):
        out = torch.empty((n, 6), device="cuda")
        tbl.SetVariant(variant)
        assert osd.B200Evaluator.EvalStencils(src, D(0, 6, 6), out, D(0, 6, 6), tbl)
        tbl.SetVariant(0)
        return out
    ex, ey = ev(x), ev(y)
    # partition of unity: refinement weights are convex, a constant field is reproduced
    const = torch.full((ncv, 6), 2.5, device="cuda")
    assert (ev(const) - 2.5).abs().max().item() <= 2.5 * REL_TOL
    # convexity: every refined value lies inside the control-value range
    assert ex.max() <= x.max() + 1e-5 and ex.min() >= x.min() - 1e-5
    # linearity
    lin = ev(0.75 * x - 1.5 * y)
    assert (lin - (0.75 * ex - 1.5 * ey)).abs().max().item() <= 2e-6 * max(1.0, lin.abs().max().item())
    # every kernel variant and the raw reference-layout path agree
    for v in (1, 2, 8, 11, 12):
        assert (ev(x, v) - ex).abs().max().item() <= 2e-6
    raw = torch.empty((n, 6), device="cuda")
    assert osd.B200Evaluator.EvalStencilsRaw(x, D(0, 6, 6), [(raw, D(0, 6, 6))], tbl.GetSizesBuffer(), tbl.GetOffsetsBuffer(),
                                             tbl.GetIndicesBuffer(), [tbl.GetWeightsBuffer()], 0, n)
    assert (raw - ex).abs().max().item() <= 2e-6
    # batched instances: 8 control-point sets through one pass over the table
    B, L3 = 8, 3
    xs = torch.randn((B, ncv, L3), device="cuda", generator=g)
    outb = torch.empty((B, n, L3), device="cuda")
    one = torch.empty((n, L3), device="cuda")
    assert osd.B200Evaluator.EvalStencilsBatched(xs, D(0, L3, L3), outb, D(0, L3, L3), tbl, B, ncv * L3, n * L3)
    assert osd.B200Evaluator.EvalStencils(xs[5], D(0, L3, L3), one, D(0, L3, L3), tbl)
    assert torch.equal(outb[5], one)
    for fn, label in ((lambda: osd.B200Evaluator.EvalStencilsBatched(xs, D(0, L3, L3), outb, D(0, L3, L3), tbl, B, ncv * L3, n * L3), "batched x8"),
                      (lambda: [osd.B200Evaluator.EvalStencils(xs[i], D(0, L3, L3), outb[i], D(0, L3, L3), tbl) for i in range(B)], "8 calls")):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"BATCHED-INSTANCES cfg2 L=3 {label}: {e0.elapsed_time(e1) / 20:.4f} ms for {B} instances "
              f"({B * n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9:.1f} G verts/s)")
    # independent cross-check of the whole 6.4 M-row result: torch sparse CSR matmul of the same table
    crow = torch.from_numpy(np.concatenate([table.offsets.astype(np.int64), [table.num_elements]])).cuda()
    A = torch.sparse_csr_tensor(crow, torch.from_numpy(table.indices.astype(np.int64)).cuda(),
                                torch.from_numpy(table.weights).cuda(), size=(n, ncv))
    assert (A @ x - ex).abs().max().item() <= 5e-6


def test_unfactorized_table_applied_level_by_level():
    """factorizeIntermediateLevels = false (far/stencilTableFactory.h:66-75, far tutorial 4_3): the rows of level l index
    the vertices of level l-1, so the table is applied one level at a time with the absolute row range -- src = the
    previous level's block of the buffer -- and gives the factorized table's result.  All levels in ONE call on the
    Osd::Mesh::Refine layout would read rows it is writing: refused (B200OSD_ERR_UNSUPPORTED), not raced."""
    from oracle import ref
    from opensubdiv_b200 import capi
    if not ref.available():
        pytest.skip("oracle/_ref/libosdref.so not present")
    for shape, level in (("catmark_cube_creases0", 4), ("catmark_car", 3), ("loop_icosahedron", 3)):
        m = ref.Mesh.from_shape(shape).refine_uniform(level)
        fact = m.stencil_table(intermediate_levels=True, factorize=True)
        unf = m.stencil_table(intermediate_levels=True, factorize=False)
        ncv, n = unf.num_control_verts, unf.num_stencils
        assert unf.indices.max() >= ncv and fact.indices.max() < ncv
        tbl = osd.B200StencilTable.Create(unf)
        assert not tbl.IsFactorized() and osd.B200StencilTable.Create(fact).IsFactorized()
        want = oracle_stencils(m.positions, (0, 3, 3), n, 3, fact, 1)[0]
        scale = oracle_stencils(m.positions, (0, 3, 3), n, 3, fact, 1, abs_scale=True)[0]
        for v in (0, 1, 8):
            tbl.SetVariant(v)
            vb = osd.B200VertexBuffer.Create(3, ncv + n)
            vb.UpdateData(np.ascontiguousarray(m.positions, np.float32), 0, ncv)
            src_vertex, row = 0, 0                               # level l-1 starts at vertex src_vertex; level l's first row
            for lv in range(1, level + 1):
                nv = m.level_num_verts(lv)
                # row i of the table is written to element i of dst (absolute addressing): dst = the refined region
                assert osd.B200Evaluator.EvalStencils(vb, D(src_vertex * 3, 3, 3), vb, D(ncv * 3, 3, 3), tbl, start=row, end=row + nv)
                src_vertex = ncv + row
                row += nv
            assert row == n
            osd.B200Evaluator.Synchronize()
            got = vb.as_tensor()[ncv:].cpu().numpy()
            # level-by-level interpolation accumulates rounding over `level` steps
            assert_close(got, want, scale, f"{shape} unfactorized, level by level, variant {v}", tol=1e-6 * level)
        tbl.SetVariant(0)
        vb = osd.B200VertexBuffer.Create(3, ncv + n)
        with pytest.raises(capi.B200OsdError, match="one level at a time"):
            osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), tbl)


def test_non_finite_control_vertex_reaches_only_the_rows_that_reference_it():
    """padded slots of the bucketed layout carry weight 0 -- they must not pull a NaN / Inf of some other vertex
    into a row (0 * NaN = NaN).  The reference only propagates a non-finite value to rows that reference the vertex."""
    d = golden("stencils_catmark_car")
    t = table_from(d, "t_")
    n = t.num_stencils
    src = np.ascontiguousarray(d["src"], np.float32).copy()
    bad = int(t.indices[t.offsets[n // 3]])                     # a vertex some rows reference
    for poison in (np.nan, np.inf):
        src[bad] = poison
        rowid = np.repeat(np.arange(n), t.sizes)
        touched = np.zeros(n, bool)
        touched[rowid[t.indices == bad]] = True
        assert touched.any() and not touched.all()
        for idx16 in (True, False):
            for kw in (dict(), dict(keep_order=True)):
                tbl = osd.B200StencilTable.Create(t, idx16=idx16, **kw)
                for v in (0, 1, 8, 122):
                    out = torch.zeros((n, 3), device="cuda")
                    tbl.SetVariant(v)
                    assert osd.B200Evaluator.EvalStencils(dev(src), D(0, 3, 3), out, D(0, 3, 3), tbl)
                    got = out.cpu().numpy()
                    assert np.isfinite(got[~touched]).all(), f"poison {poison} leaked (idx16={idx16}, {kw}, variant {v})"
                    assert (~np.isfinite(got[touched])).all(axis=1).all()
