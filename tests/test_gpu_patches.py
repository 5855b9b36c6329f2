"""EvalPatches / EvalPatchesVarying / EvalPatchesFaceVarying parity on a real B200 (through the C ABI) vs the oracle
and the reference's golden outputs."""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from opensubdiv_b200 import synth
from tests.gpu_util import D, dev, coords_dev, oracle_patches, oracle_stencils
from tests.util import golden, golden_names, table_from, triple_from, assert_close, assert_close_bbox, REL_TOL

pytestmark = pytest.mark.gpu
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")


class _PT:
    def __init__(self, vertex, varying=None, fvar=None):
        self.vertex, self.varying, self.fvar = vertex, varying, fvar or []


@pytest.mark.parametrize("name", golden_names("patches_"))
def test_golden_patch_tables_all_arities(name):
    d = golden(name)
    vtx = triple_from(d, "vtx_")
    var = triple_from(d, "var_") if "var_arrays" in d.files else None
    fv = triple_from(d, "fvar_") if "fvar_arrays" in d.files else None
    pt = osd.B200PatchTable.Create(_PT(vtx, var, [fv] if fv is not None else []))
    assert pt is not None
    coords = d["coords"]
    n = len(coords)
    pc = coords_dev(coords)
    src = dev(d["vb"])
    scales = oracle_patches(d["vb"], (0, 3, 3), 3, coords, vtx, 6, abs_scale=True)
    # second gate, free of any weight-derived scale: error relative to bbox * 2^(order * depth).  P and 1st derivatives
    # meet 1e-6; 2nd derivatives are formed from weights of size 4^depth that cancel against each other and differ from the
    # reference's evaluation order by up to 1.7e-6 of that scale on boundary / triangle fixtures (the reference's own fp32
    # result is as far from the double-precision value: tests/test_accuracy_vs_f64.py), so their gate is 2.5e-6.
    bbox = float((d["vb"].max(axis=0) - d["vb"].min(axis=0)).max())
    depth = (vtx.params["field1"][coords["patchIndex"]] & 0xF).astype(np.int64)
    order_of = (0, 1, 1, 2, 2, 2)
    for variant in (0, 1, 2, 3):    # automatic, caller's order, grouped by patch on the device, per-call hull cache
      for nw in (1, 3, 6):
        # outputs interleaved in one buffer, glEvalLimit style (examples/glEvalLimit/glEvalLimit.cpp:277-287)
        out = torch.full((n, 3 * nw), float("nan"), device="cuda")
        args = []
        for k in range(nw):
            args += [out, D(3 * k, 3, 3 * nw)]
        pt.SetVariant(variant)
        try:
            assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, None)
        finally:
            pt.SetVariant(0)
        res = out.cpu().numpy()
        for k in range(nw):
            assert_close(res[:, 3 * k:3 * k + 3], d["out_" + OUT6[k]], scales[k], f"{name} variant={variant} nw={nw} {OUT6[k]}")
            assert_close_bbox(res[:, 3 * k:3 * k + 3], d["out_" + OUT6[k]], bbox, depth, order_of[k],
                              f"{name} variant={variant} nw={nw} {OUT6[k]} (bbox gate)", tol=1e-6 if order_of[k] < 2 else 2.5e-6)
    # raw-pointer overload (reference-layout device arrays, osd/cudaEvaluator.h:815-827)
    outs = [torch.zeros((n, 3), device="cuda") for _ in range(6)]
    assert osd.B200Evaluator.EvalPatchesRaw(src, D(0, 3, 3), [(o, D(0, 3, 3)) for o in outs], n, pc, pt.GetPatchArrayBuffer(),
                                            pt.GetPatchIndexBuffer(), pt.GetPatchParamBuffer())
    for k in range(6):
        assert_close(outs[k].cpu().numpy(), d["out_" + OUT6[k]], scales[k], f"{name} raw {OUT6[k]}")
    if var is not None:
        vsrc = dev(d["var_vb"])
        vs = oracle_patches(d["var_vb"], (0, 3, 3), 3, coords, var, 3, abs_scale=True)
        outs = [torch.zeros((n, 3), device="cuda") for _ in range(3)]
        args = []
        for o in outs:
            args += [o, D(0, 3, 3)]
        assert osd.B200Evaluator.EvalPatchesVarying(vsrc, D(0, 3, 3), *args, n, pc, pt, None)
        for k in range(3):
            assert_close(outs[k].cpu().numpy(), d["var_out_" + OUT6[k]], vs[k], f"{name} varying {OUT6[k]}")
    if fv is not None:
        fsrc = dev(d["fvar_vb"])
        fs = oracle_patches(d["fvar_vb"], (0, 2, 2), 2, coords, fv, 6, abs_scale=True)
        outs = [torch.zeros((n, 2), device="cuda") for _ in range(6)]
        args = []
        for o in outs:
            args += [o, D(0, 2, 2)]
        for variant in (1, 2, 3):
            pt.SetVariant(variant)
            try:
                assert osd.B200Evaluator.EvalPatchesFaceVarying(fsrc, D(0, 2, 2), *args, n, pc, pt, 0, None)
            finally:
                pt.SetVariant(0)
            for k in range(6):
                assert_close(outs[k].cpu().numpy(), d["fvar_out_" + OUT6[k]], fs[k], f"{name} fvar v{variant} {OUT6[k]}")


def test_refine_then_evaluate_pipeline():
    """The glEvalLimit frame: UpdateData(cv) -> EvalStencils incl. end-cap local points -> EvalPatches (glEvalLimit.cpp:371-449)."""
    d = golden("patches_catmark_car")
    st = table_from(d, "st_")
    vtx = triple_from(d, "vtx_")
    ncv, n = st.num_control_verts, st.num_stencils
    vb = osd.B200VertexBuffer.Create(3, ncv + n)
    vb.UpdateData(np.ascontiguousarray(d["src0"]), 0, ncv)
    stbl = osd.B200StencilTable.Create(st)
    assert osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), stbl)
    pt = osd.B200PatchTable.Create(_PT(vtx))
    coords = d["coords"]
    out = osd.B200VertexBuffer.Create(18, len(coords))
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, len(coords), coords_dev(coords), pt, None)
    osd.B200Evaluator.Synchronize()
    res = np.zeros((len(coords), 18), np.float32)
    out.ReadData(res, 0, len(coords))
    osd.B200Evaluator.Synchronize()
    scales = oracle_patches(d["vb"], (0, 3, 3), 3, coords, vtx, 6, abs_scale=True)
    for k in range(6):
        assert_close(res[:, 3 * k:3 * k + 3], d["out_" + OUT6[k]], scales[k] * 1.5 + 1e-6, f"pipeline {OUT6[k]}")


@pytest.mark.parametrize("L,stride,offset", [(1, 1, 0), (2, 2, 0), (4, 4, 0), (6, 6, 0), (3, 5, 2), (9, 12, 1)])
def test_primvar_lengths_and_null_outputs(L, stride, offset):
    d = golden("patches_catmark_gregory_test2")
    vtx = triple_from(d, "vtx_")
    coords = d["coords"]
    n = len(coords)
    nv = len(d["vb"])
    rng = np.random.default_rng(L)
    src = rng.standard_normal(offset + nv * stride + 3).astype(np.float32)
    exp = oracle_patches(src, (offset, L, stride), L, coords, vtx, 6)
    scl = oracle_patches(src, (offset, L, stride), L, coords, vtx, 6, abs_scale=True)
    pt = osd.B200PatchTable.Create(_PT(vtx))
    for variant in (1, 2, 3):
        outs = [torch.full((n, L), float("nan"), device="cuda") for _ in range(6)]
        # NULL du and dvv are skipped (osd/cudaKernel.cu:300-327)
        bufs = [outs[0], None, outs[2], outs[3], outs[4], None]
        args = []
        for b in bufs:
            args += [b, D(0, L, L)]
        pt.SetVariant(variant)
        try:
            assert osd.B200Evaluator.EvalPatches(dev(src), D(offset, L, stride), *args, n, coords_dev(coords), pt, None)
        finally:
            pt.SetVariant(0)
        for k in (0, 2, 3, 4):
            assert_close(outs[k].cpu().numpy(), exp[k], scl[k], f"L={L} v{variant} {OUT6[k]}")
        assert torch.isnan(outs[1]).all() and torch.isnan(outs[5]).all()
    # errors: NULL src, value-only NULL dst, length mismatch -> false (osd/cpuEvaluator.cpp:165-176)
    assert not osd.B200Evaluator.EvalPatches(None, D(0, L, L), outs[0], D(0, L, L), n, coords_dev(coords), pt, None)
    assert not osd.B200Evaluator.EvalPatches(dev(src), D(0, L, stride), None, D(0, L, L), n, coords_dev(coords), pt, None)
    assert not osd.B200Evaluator.EvalPatches(dev(src), D(0, L, stride), outs[0], D(0, L + 1, L + 1), n, coords_dev(coords), pt, None)
    assert osd.B200Evaluator.EvalPatches(dev(src), D(0, L, stride), outs[0], D(0, L, L), 0, coords_dev(coords), pt, None)   # empty


@pytest.fixture(scope="module")
def torus_patches():
    mesh = synth.torus_quads(400, 250)
    ptab = synth.torus_patch_table(mesh)
    pt = osd.B200PatchTable.Create(ptab)
    return mesh, ptab, pt


@pytest.mark.slow
def test_ten_million_random_coords_vs_oracle_and_order_independence(torus_patches):
    """Config-4-sized run: 10 M random PatchCoords, P + 1st + 2nd derivatives."""
    mesh, ptab, pt = torus_patches
    n = 10_000_000
    coords = synth.random_patch_coords(len(mesh.faces), n, seed=2024)
    src_np = synth.deform(mesh.positions, 3)
    src = dev(src_np)
    pc = coords_dev(coords)
    out = torch.empty((n, 18), device="cuda")
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, None)
    # oracle on a bounded sample of the same coords
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(n, 200_000, replace=False))
    exp = oracle_patches(src_np, (0, 3, 3), 3, coords[pick], ptab.vertex, 6)
    # all patches of the torus are interior regular B-splines: the plain scale sum|w||x| is the right yardstick
    scl = oracle_patches(src_np, (0, 3, 3), 3, coords[pick], ptab.vertex, 6, abs_scale=True, plain=True)
    got = out[torch.from_numpy(pick).cuda()].cpu().numpy()
    for k in range(6):
        assert_close(got[:, 3 * k:3 * k + 3], exp[k], scl[k], f"10M {OUT6[k]}")
    # the same coords sorted by patch must give identical bits per coordinate
    order = np.argsort(coords["patchIndex"], kind="stable")
    out2 = torch.empty((n, 18), device="cuda")
    args2 = []
    for k in range(6):
        args2 += [out2, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args2, n, coords_dev(coords[order]), pt, None)
    assert torch.equal(out2, out[torch.from_numpy(order).cuda()])
    # partition of unity: constant field -> P = const, all derivatives ~ 0
    const = torch.full((mesh.num_verts, 3), 1.25, device="cuda")
    assert osd.B200Evaluator.EvalPatches(const, D(0, 3, 3), *args, n, pc, pt, None)
    assert (out[:, 0:3] - 1.25).abs().max().item() <= 1.25 * 2 * REL_TOL
    assert out[:, 3:].abs().max().item() <= 1e-5


@pytest.mark.slow
def test_limit_stencils_agree_with_patch_evaluation_on_gpu(torus_patches):
    """Two independent GPU paths that must agree: LimitStencilTable rows (EvalStencils, 6 weight streams) and
    EvalPatches at the same (face, s, t)."""
    mesh, ptab, pt = torus_patches
    n = 1_000_000
    rng = np.random.default_rng(12345)
    face = np.sort(rng.integers(0, len(mesh.faces), n)).astype(np.int32)
    s, t = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    ls = synth.torus_limit_stencil_table(mesh, face, s, t)
    stbl = osd.B200StencilTable.Create(ls)
    src = dev(mesh.positions)
    a = torch.empty((n, 18), device="cuda")
    b = torch.empty((n, 18), device="cuda")
    args_a, args_b = [], []
    for k in range(6):
        args_a += [a, D(3 * k, 3, 18)]
        args_b += [b, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), *args_a, stbl)
    coords = np.zeros(n, osd.PATCH_COORD_DTYPE)
    coords["patchIndex"] = face
    coords["vertIndex"] = face * 16
    coords["s"], coords["t"] = s, t
    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args_b, n, coords_dev(coords), pt, None)
    assert (a - b).abs().max().item() <= 2e-5
    # and the stencil side against the oracle on a window
    exp = oracle_stencils(mesh.positions, (0, 3, 3), n, 3, ls, 6, 0, 20000)
    scl = oracle_stencils(mesh.positions, (0, 3, 3), n, 3, ls, 6, 0, 20000, abs_scale=True)
    got = a[:20000].cpu().numpy()
    for k in range(6):
        assert_close(got[:, 3 * k:3 * k + 3], exp[k][:20000], scl[k][:20000], f"limit {OUT6[k]}")


@pytest.mark.slow
def test_config4_adaptive_gregory_tables_ten_million_coords():
    """BASELINE config 4 at full size with REAL Far tables: 60 tiled copies of regression shape catmark_car (98 520
    control vertices, creases, extraordinary vertices), adaptive level 3 with Gregory-basis end caps -> 1.24 M REGULAR +
    75 k GREGORY_BASIS patches, local-point stencils appended; 10 M random PatchCoords from Far::PatchMap::FindPatch;
    P + 1st + 2nd derivatives (interleaved like glEvalLimit) and face-varying UVs; checked against the oracle on a sample."""
    import time
    from oracle import oracle, ref
    if not ref.available():
        pytest.skip("oracle/_ref/libosdref.so not present")
    m = ref.Mesh.from_shape_tiled("catmark_car", 60)
    ptab = m.patch_table(3, end_cap="gregory", fvar=True, fvar_legacy_linear=False, inf_sharp=True, legacy_sharp_corner=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=ptab)
    fst = m.stencil_table(mode="fvar", intermediate_levels=True, patch_table=ptab)
    assert set(int(x) for x in ptab.vertex.arrays["desc"]) == {6, 9}
    n = 10_000_000
    rng = np.random.default_rng(2024)
    face = rng.integers(0, m.num_ptex_faces, n).astype(np.int32)
    s, t = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    coords = m.find_patches(ptab, face, s, t)
    assert (coords["arrayIndex"] >= 0).all()

    ncv, nst = st.num_control_verts, st.num_stencils
    vb = osd.B200VertexBuffer.Create(3, ncv + nst)
    vb.UpdateData(np.ascontiguousarray(m.positions), 0, ncv)
    stbl = osd.B200StencilTable.Create(st)
    pt = osd.B200PatchTable.Create(ptab)

    # the same 10 M samples located on the device -- every record bit-identical to Far::PatchMap::FindPatch
    pm = osd.B200PatchMap.Create(ptab)
    dface, ds, dt = dev(face), dev(s), dev(t)
    pc = torch.zeros(n * 5, dtype=torch.int32, device="cuda")
    found = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert pm.FindPatches(n, dface, ds, dt, pc, found)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        pm.FindPatches(n, dface, ds, dt, pc, found)
    f1.record()
    torch.cuda.synchronize()
    print(f"CONFIG4 FindPatches(10M samples, depth<={pm.GetMaxDepth()}, {pm.GetNumNodes()} nodes): "
          f"{f0.elapsed_time(f1) / 5:.3f} ms")
    assert int(found.item()) == n
    assert np.array_equal(pc.cpu().numpy(), np.ascontiguousarray(coords).view(np.int32))
    out = torch.empty((n, 18), device="cuda")
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]

    def frame():
        assert osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), stbl)
        assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, None)
    frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        frame()
    e1.record()
    torch.cuda.synchronize()
    print(f"CONFIG4 refine({nst} stencils)+EvalPatches(10M coords, 6 outputs, {len(ptab.vertex.params)} patches): "
          f"{e0.elapsed_time(e1) / 5:.3f} ms/frame")

    # oracle: refine on the CPU, evaluate a sample of the same coordinates
    cpu_vb = np.zeros((ncv + nst, 3), np.float32)
    cpu_vb[:ncv] = m.positions
    assert oracle.eval_stencils(cpu_vb.reshape(-1), (0, 3, 3), [cpu_vb.reshape(-1)], [(ncv * 3, 3, 3)], st.sizes, st.offsets,
                                st.indices, [st.weights])
    pick = np.sort(rng.choice(n, 200_000, replace=False))
    # make sure the sample contains Gregory patches
    greg = np.nonzero(coords["arrayIndex"] == 1)[0][:20_000]
    pick = np.unique(np.concatenate([pick, greg]))
    sel = np.ascontiguousarray(coords[pick])
    exp = oracle_patches(cpu_vb, (0, 3, 3), 3, sel, ptab.vertex, 6)
    scl = oracle_patches(cpu_vb, (0, 3, 3), 3, sel, ptab.vertex, 6, abs_scale=True)
    got = out[torch.from_numpy(pick).cuda()].cpu().numpy()
    for k in range(6):
        assert_close(got[:, 3 * k:3 * k + 3], exp[k], scl[k], f"config4 {OUT6[k]}")

    # face-varying UVs through EvalPatchesFaceVarying(channel 0) on the refined fvar buffer
    nfv, nfst = fst.num_control_verts, fst.num_stencils
    fvb = osd.B200VertexBuffer.Create(2, nfv + nfst)
    fvb.UpdateData(np.ascontiguousarray(m.uvs[:nfv]), 0, nfv)
    ftbl = osd.B200StencilTable.Create(fst)
    assert osd.B200Evaluator.EvalStencils(fvb, D(0, 2, 2), fvb, D(nfv * 2, 2, 2), ftbl)
    uv = torch.empty((n, 2), device="cuda")
    assert osd.B200Evaluator.EvalPatchesFaceVarying(fvb, D(0, 2, 2), uv, D(0, 2, 2), n, pc, pt, 0, None)
    cpu_f = np.zeros((nfv + nfst, 2), np.float32)
    cpu_f[:nfv] = m.uvs[:nfv]
    oracle.eval_stencils(cpu_f.reshape(-1), (0, 2, 2), [cpu_f.reshape(-1)], [(nfv * 2, 2, 2)], fst.sizes, fst.offsets, fst.indices,
                         [fst.weights])
    fexp = oracle_patches(cpu_f, (0, 2, 2), 2, sel, ptab.fvar[0], 1)
    fscl = oracle_patches(cpu_f, (0, 2, 2), 2, sel, ptab.fvar[0], 1, abs_scale=True)
    assert_close(uv[torch.from_numpy(pick).cuda()].cpu().numpy(), fexp[0], fscl[0], "config4 fvar uv")

    # sorted-vs-shuffled coordinate order gives identical bits per coordinate
    order = np.argsort(coords["patchIndex"], kind="stable")
    out2 = torch.empty((n, 18), device="cuda")
    args2 = []
    for k in range(6):
        args2 += [out2, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args2, n, coords_dev(coords[order]), pt, None)
    assert torch.equal(out2, out[torch.from_numpy(order).cuda()])


# ------------------------------------------------------------------ device-side grouping by patch --
def _eval18(src, n, pc, pt, instance=None, ctx=None, fill=float("nan")):
    out = torch.full((n, 18), fill, device="cuda")
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, instance, ctx)
    return out


@pytest.mark.parametrize("n", [70_000, 300_001])
def test_grouping_by_patch_is_bit_identical_per_index(n):
    """random coordinates are grouped by patch on the device (counting sort) and every result is
    written at the CALLER's index.  Never / automatic / always grouping, the cached plan of an evaluator instance, holes
    (arrayIndex = -1, outputs untouched) and an already sorted set (left alone by the probe) must agree bit for bit."""
    mesh = synth.torus_quads(60, 40)
    ptab = synth.torus_patch_table(mesh)
    pt = osd.B200PatchTable.Create(ptab)
    coords = synth.random_patch_coords(len(mesh.faces), n, seed=n)
    holes = np.arange(7, n, 1013)
    coords["arrayIndex"][holes] = -1
    src = dev(synth.deform(mesh.positions, 2))
    pc = coords_dev(coords)
    pt.SetVariant(1)
    ref = _eval18(src, n, pc, pt)
    live = torch.ones(n, dtype=torch.bool, device="cuda")
    live[torch.from_numpy(holes).cuda()] = False
    assert torch.isnan(ref[~live]).all() and not torch.isnan(ref[live]).any()
    keep = np.nonzero(coords["arrayIndex"][:5000] >= 0)[0]          # the oracle, like the reference, knows no "no patch" records
    sel = np.ascontiguousarray(coords[keep])
    exp = oracle_patches(synth.deform(mesh.positions, 2), (0, 3, 3), 3, sel, ptab.vertex, 6)
    scl = oracle_patches(synth.deform(mesh.positions, 2), (0, 3, 3), 3, sel, ptab.vertex, 6, abs_scale=True)
    got = ref[:5000].cpu().numpy()
    for k in range(6):
        assert_close(got[keep, 3 * k:3 * k + 3], exp[k], scl[k], f"ungrouped {OUT6[k]}")
    for variant in (0, 2, 3):
        pt.SetVariant(variant)
        got = _eval18(src, n, pc, pt)
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)), f"variant {variant}"
    pt.SetVariant(0)
    # the instantiatable evaluator caches the grouping of one coordinate set (osd/mesh.h:305-409)
    inst = osd.B200Evaluator.Create(D(0, 3, 3), D(0, 3, 18))
    assert inst.BindPatchCoords(n, pc, pt)
    for _ in range(2):
        got = _eval18(src, n, pc, pt, instance=inst)
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)), "cached plan"
    # separate (non-interleaved) outputs through the cached grouping
    outs = [torch.full((n, 3), float("nan"), device="cuda") for _ in range(3)]
    a = []
    for o in outs:
        a += [o, D(0, 3, 3)]
    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *a, n, pc, pt, inst)
    for k in range(3):
        assert torch.equal(torch.nan_to_num(outs[k], nan=-7.0), torch.nan_to_num(ref[:, 3 * k:3 * k + 3], nan=-7.0)), f"separate {k}"
    # a different coordinate buffer does not match the bound set: falls back to the per-call path, same bits
    pc2 = pc.clone()
    got = _eval18(src, n, pc2, pt, instance=inst)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
    # already coherent coordinates: the probe leaves them in place
    order = np.argsort(coords["patchIndex"], kind="stable")
    got = _eval18(src, n, coords_dev(coords[order]), pt)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref[torch.from_numpy(order).cuda()], nan=-7.0))


def test_one_table_two_threads_two_streams():
    """VERDICT r1 / ADVICE: the table is immutable and evaluation scratch is per call and stream-ordered, so two host
    threads may evaluate ONE table on two streams at the same time."""
    import threading
    mesh = synth.torus_quads(60, 40)
    ptab = synth.torus_patch_table(mesh)
    pt = osd.B200PatchTable.Create(ptab)
    n = 200_000
    src = dev(synth.deform(mesh.positions, 1))
    sets = [coords_dev(synth.random_patch_coords(len(mesh.faces), n, seed=s)) for s in (1, 2)]
    pt.SetVariant(1)
    want = [_eval18(src, n, pc, pt) for pc in sets]
    pt.SetVariant(0)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [None, None]
    errs = []

    def work(k):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(streams[k]):          # thread-local current stream: fills and kernels are ordered on it
                for _ in range(20):
                    got[k] = _eval18(src, n, sets[k], pt)
            streams[k].synchronize()
        except Exception as exc:      # surfaced in the main thread
            errs.append(exc)
    th = [threading.Thread(target=work, args=(k,)) for k in (0, 1)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for k in (0, 1):
        assert torch.equal(got[k], want[k])


@pytest.mark.parametrize("ptype", [3, 4, 5, 6, 9, 10])
def test_every_boundary_mask_depth_rotation_on_the_gpu(ptype):
    """The synthetic PatchParam sweep of tests/test_kernel_math_emu.py (all 32 Loop / 16 B-spline boundary masks, depths
    0..6, rotated triangles, corner and edge samples) through the real kernels in every serving mode, against the oracle."""
    from tests.test_kernel_math_emu import synthetic_patch_sweep
    tr, coords, vb = synthetic_patch_sweep(ptype)
    n = len(coords)
    exp = oracle_patches(vb, (0, 3, 3), 3, coords, tr, 6)
    scl = oracle_patches(vb, (0, 3, 3), 3, coords, tr, 6, abs_scale=True)
    pt = osd.B200PatchTable.Create(_PT(tr))
    pc, src = coords_dev(coords), dev(vb)
    first = None
    for variant in (1, 2, 3):
        pt.SetVariant(variant)
        out = torch.full((n, 18), float("nan"), device="cuda")
        args = []
        for k in range(6):
            args += [out, D(3 * k, 3, 18)]
        assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, None)
        res = out.cpu().numpy()
        if first is None:
            first = res
            for k in range(6):
                assert_close(res[:, 3 * k:3 * k + 3], exp[k], scl[k], f"type {ptype} {OUT6[k]}")
        else:
            assert np.array_equal(res, first), f"type {ptype}: variant {variant} differs bitwise"
