"""The C oracle against the unmodified reference compiled in place (oracle/_ref/libosdref.so), on inputs beyond the
committed fixtures.  Skipped where the reference build is absent."""
import numpy as np
import pytest

from oracle import oracle, ref
from tests.util import assert_close

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libosdref.so not built (needs /root/reference)")


def _f1(depth, nonquad, regular, boundary, v, u):
    return depth | (nonquad << 4) | (regular << 5) | (boundary << 7) | (v << 12) | (u << 22)


@pytest.mark.parametrize("ptype", [3, 4, 5, 6, 9, 10])
def test_patch_basis_all_types_masks_depths(ptype):
    """OsdEvaluatePatchBasis (osd/patchBasis.h:1555-1610) for every basis, boundary mask, depth, rotation."""
    rng = np.random.default_rng(ptype)
    tri = ptype in (4, 5, 10)
    masks = range(32) if ptype == 5 else (range(16) if ptype == 6 else [0])
    for mask in masks:
        for trial in range(60):
            depth = int(rng.integers(0, 7))
            nonquad = int(rng.integers(0, 2)) if depth > 0 else 0
            n = 1 << (depth - nonquad)
            u, v = int(rng.integers(0, n)), int(rng.integers(0, n))
            a, b = rng.random(2)
            if tri and a + b > 1:
                a, b = 1 - a, 1 - b
            if trial % 7 == 0:
                a = 0.0
            if trial % 11 == 0:
                b = 0.0
            if trial % 13 == 0:
                a, b = (1.0, 0.0) if tri else (1.0, 1.0)
            rotated = tri and (u + v) >= (1 << depth)
            if rotated:
                s, t = ((1 << depth) - u - a) / n, ((1 << depth) - v - b) / n
            else:
                s, t = (u + a) / n, (v + b) / n
            f1 = _f1(depth, nonquad, 1, mask, v, u)
            for nw in (1, 3, 6):
                nr, wr = ref.osd_patch_basis(ptype, 0, f1, np.float32(s), np.float32(t), nw)
                no, wo = oracle.patch_basis(ptype, 0, f1, np.float32(s), np.float32(t), nw)
                assert nr == no
                for k in range(nw):
                    if ptype == 10:     # Bezier triangle restated in Bernstein form: agrees to ~1 ulp of the weight scale
                        sc = np.abs(wr[k][:nr]).max()
                        assert np.abs(wr[k][:nr] - wo[k][:nr]).max() <= 1e-6 * max(sc, 1e-30)
                    else:
                        assert np.array_equal(wr[k][:nr], wo[k][:nr]), (ptype, mask, depth, k)


@pytest.mark.parametrize("shape,level", [("catmark_cube", 4), ("catmark_helmet", 2), ("loop_saddle_edgecorner", 3),
                                         ("catmark_pole360", 1), ("catmark_chaikin0", 3), ("catmark_nonquads", 3)])
@pytest.mark.parametrize("L,stride,offset", [(3, 3, 0), (4, 4, 0), (8, 8, 0), (6, 9, 2), (1, 5, 4)])
def test_eval_stencils_descriptors(shape, level, L, stride, offset):
    m = ref.Mesh.from_shape(shape).refine_uniform(level)
    st = m.stencil_table(intermediate_levels=True)
    rng = np.random.default_rng(3)
    ncv, n = st.num_control_verts, st.num_stencils
    src = rng.standard_normal(offset + ncv * stride + 8).astype(np.float32)
    a = np.full(n * stride + offset + 8, np.nan, np.float32)
    b = a.copy()
    assert ref.eval_stencils(src, (offset, L, stride), [a], [(offset, L, stride)], st)
    assert oracle.eval_stencils(src, (offset, L, stride), [b], [(offset, L, stride)], st.sizes, st.offsets, st.indices,
                                [st.weights])
    assert np.array_equal(a, b, equal_nan=True)


def test_far_update_values_cross_check():
    """Third independent implementation: Far::StencilTable::UpdateValues (far/stencilTable.h:648-674)."""
    m = ref.Mesh.from_shape("catmark_bishop").refine_uniform(2)
    st = m.stencil_table()
    pos = m.positions
    far = ref.far_update_values_xyz(m, pos)
    out = np.zeros((st.num_stencils, 3), np.float32)
    assert oracle.eval_stencils(pos.reshape(-1), (0, 3, 3), [out.reshape(-1)], [(0, 3, 3)], st.sizes, st.offsets,
                                st.indices, [st.weights])
    assert np.abs(far - out).max() <= 1e-6


@pytest.mark.parametrize("shape,level,endcap", [("catmark_pawn", 3, "gregory"), ("catmark_rook", 2, "bspline"),
                                                ("loop_toroidal_tet", 3, "gregory"), ("catmark_flap", 3, "gregory"),
                                                ("catmark_single_crease", 3, "gregory")])
def test_eval_patches_live(shape, level, endcap):
    m = ref.Mesh.from_shape(shape)
    pt = m.patch_table(level, end_cap=endcap, inf_sharp=True, single_crease=(shape == "catmark_single_crease"),
                       legacy_sharp_corner=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    ncv, n = st.num_control_verts, st.num_stencils
    vb = np.zeros((ncv + n, 3), np.float32)
    vb[:ncv] = m.positions
    ref.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st)
    rng = np.random.default_rng(1)
    k = 3000
    face = rng.integers(0, m.num_ptex_faces, k).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    if m.reg_face_size == 3:
        flip = s + t >= 1
        s, t = np.where(flip, 1 - s, s).astype(np.float32), np.where(flip, 1 - t, t).astype(np.float32)
    pc = m.find_patches(pt, face, s, t)
    pc = pc[pc["arrayIndex"] >= 0]
    for nw in (1, 3, 6):
        a = [np.zeros((len(pc), 3), np.float32) for _ in range(nw)]
        b = [np.zeros((len(pc), 3), np.float32) for _ in range(nw)]
        sc = [np.zeros((len(pc), 3), np.float32) for _ in range(nw)]
        assert ref.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in a], [(0, 3, 3)] * nw, pc, pt.vertex)
        args = (pc, pt.vertex.arrays, pt.vertex.indices, pt.vertex.params)
        assert oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in b], [(0, 3, 3)] * nw, *args)
        with oracle.abs_mode(2):
            oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in sc], [(0, 3, 3)] * nw, *args)
        for x, y, z in zip(a, b, sc):
            assert_close(y, x, z, shape)


def test_limit_stencils_agree_with_patch_evaluation():
    """Two reference code paths that must agree: LimitStencilTable applied to the control cage vs
    EvalPatches on the refined buffer, at the same locations."""
    m = ref.Mesh.from_shape("catmark_cube_creases1")
    pt = m.patch_table(3, end_cap="gregory")
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    rng = np.random.default_rng(9)
    k = 500
    face = np.sort(rng.integers(0, m.num_ptex_faces, k)).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    ls = m.limit_stencil_table(face, s, t, True, True, patch_table=pt)
    pos = m.positions
    lim = [np.zeros((k, 3), np.float32) for _ in range(6)]
    assert oracle.eval_stencils(pos.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in lim], [(0, 3, 3)] * 6, ls.sizes,
                                ls.offsets, ls.indices, ls.weight_streams(6))
    ncv, n = st.num_control_verts, st.num_stencils
    vb = np.zeros((ncv + n, 3), np.float32)
    vb[:ncv] = pos
    oracle.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st.sizes, st.offsets, st.indices,
                         [st.weights])
    pc = m.find_patches(pt, face, s, t)
    ev = [np.zeros((k, 3), np.float32) for _ in range(6)]
    assert oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in ev], [(0, 3, 3)] * 6, pc,
                               pt.vertex.arrays, pt.vertex.indices, pt.vertex.params)
    for a, b, tol in zip(lim, ev, (2e-6, 2e-5, 2e-5, 5e-4, 5e-4, 5e-4)):
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


ALL_SHAPES = ref.shape_names() if ref.available() else []


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_every_regression_shape_stencils_and_patches(shape):
    """The reference's osd_regression walks its whole shape list (regression/osd_regression/main.cpp:250-330); so does
    the oracle: uniform level-2 stencils bit-identical, adaptive level-3 EvalPatches (Gregory end caps, 6 outputs) within
    the Gregory-triangle 3e-7, FindPatch bit-identical."""
    m = ref.Mesh.from_shape(shape).refine_uniform(2)
    st = m.stencil_table(intermediate_levels=True)
    pos = m.positions
    a, b = np.zeros((st.num_stencils, 3), np.float32), np.zeros((st.num_stencils, 3), np.float32)
    assert ref.eval_stencils(pos.reshape(-1), (0, 3, 3), [a.reshape(-1)], [(0, 3, 3)], st)
    assert oracle.eval_stencils(pos.reshape(-1), (0, 3, 3), [b.reshape(-1)], [(0, 3, 3)], st.sizes, st.offsets, st.indices,
                                [st.weights])
    assert np.array_equal(a.view(np.int32), b.view(np.int32)), shape
    if shape.startswith("bilinear"):
        return
    m = ref.Mesh.from_shape(shape)
    # the valence-360 pole costs the reference 100 s of end-cap stencil building at level 3: one level is enough there
    pt = m.patch_table(1 if shape.endswith("pole360") else 3, end_cap="gregory", inf_sharp=True, legacy_sharp_corner=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    ncv, n = st.num_control_verts, st.num_stencils
    vb = np.zeros((ncv + n, 3), np.float32)
    vb[:ncv] = m.positions
    assert ref.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st)
    rng = np.random.default_rng(3)
    k = 1500
    face = rng.integers(0, m.num_ptex_faces, k).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    if m.reg_face_size == 3:
        flip = s + t >= 1
        s, t = np.where(flip, 1 - s, s).astype(np.float32), np.where(flip, 1 - t, t).astype(np.float32)
    want = m.find_patches(pt, face, s, t)
    got = oracle.find_patches(pt.vertex.arrays, pt.vertex.params, m.reg_face_size == 3, face, s, t)
    hit = want["arrayIndex"] >= 0
    assert np.array_equal(got["arrayIndex"] >= 0, hit)
    assert np.array_equal(got[hit].view(np.int32), np.ascontiguousarray(want[hit]).view(np.int32)), shape
    pc = np.ascontiguousarray(want[hit])
    x = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    y = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    z = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    assert ref.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in x], [(0, 3, 3)] * 6, pc, pt.vertex)
    args = (pc, pt.vertex.arrays, pt.vertex.indices, pt.vertex.params)
    assert oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in y], [(0, 3, 3)] * 6, *args)
    with oracle.abs_mode(2):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in z], [(0, 3, 3)] * 6, *args)
    for xx, yy, zz in zip(x, y, z):
        assert_close(yy, xx, zz, shape, tol=3e-7)


@pytest.mark.parametrize("shape,level", [("catmark_car", 2), ("loop_icosahedron", 3), ("catmark_gregory_test2", 3),
                                         ("catmark_edgecorner", 3), ("loop_cube_creases0", 2)])
def test_far_basis_twin_equals_the_osd_mirror(shape, level):
    """Far::PatchTable::EvaluateBasis (far/patchTable.cpp:581-625, what LimitStencilTableFactory uses) and the Osd mirror
    the evaluators use (osd/patchBasis.h, restated by the oracle) are the same functions: bit-identical weights, except
    the oracle's Bernstein-form Gregory triangle (3e-7)."""
    m = ref.Mesh.from_shape(shape)
    pt = m.patch_table(level, end_cap="gregory", inf_sharp=True, legacy_sharp_corner=False)
    rng = np.random.default_rng(0)
    k = 1500
    face = rng.integers(0, m.num_ptex_faces, k).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    if m.reg_face_size == 3:
        flip = s + t >= 1
        s, t = np.where(flip, 1 - s, s).astype(np.float32), np.where(flip, 1 - t, t).astype(np.float32)
    pc = m.find_patches(pt, face, s, t)
    pc = np.ascontiguousarray(pc[pc["arrayIndex"] >= 0])
    W = ref.far_basis(pt, pc)
    for i, c in enumerate(pc):
        a, p = pt.vertex.arrays[c["arrayIndex"]], pt.vertex.params[c["patchIndex"]]
        typ = int(a["regDesc"] if (int(p["field1"]) >> 5) & 1 else a["desc"])
        n, w = oracle.patch_basis(typ, int(p["field0"]), int(p["field1"]), float(c["s"]), float(c["t"]))
        for q in range(6):
            if typ == 10:
                assert np.abs(W[q][i][:n] - w[q][:n]).max() <= 3e-7 * max(np.abs(W[q][i][:n]).max(), 1e-30)
            else:
                assert np.array_equal(W[q][i][:n].view(np.int32), w[q][:n].view(np.int32)), (shape, i, q)
