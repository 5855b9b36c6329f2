"""Limit-stencil table construction ON THE DEVICE (b200osd_limit_stencil_table_create) against
Far::LimitStencilTableFactory of the reference (oracle/_ref) and the oracle's restatement of it: same rows, same element
order, bit-identical weights for Catmark in all six streams; and the table evaluates like Far's."""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from oracle import oracle, ref
from tests.gpu_util import D, dev, coords_dev

pytestmark = pytest.mark.gpu
STREAMS = ("weights", "du", "dv", "duu", "duv", "dvv")


def locations(m, k, seed):
    rng = np.random.default_rng(seed)
    face = np.sort(rng.integers(0, m.num_ptex_faces, k)).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    s[::13], t[::17] = 0.0, 1.0                       # patch corners / edges: exact zeros among the basis weights
    s[5::29] = 0.5
    if m.reg_face_size == 3:
        flip = s + t > 1
        s, t = np.where(flip, 1 - s, s).astype(np.float32), np.where(flip, 1 - t, t).astype(np.float32)
    return face, s, t


def build(shape, level, nw, k=1500):
    if not ref.available():
        pytest.skip("oracle/_ref/libosdref.so not present")
    m = ref.Mesh.from_shape(shape)
    pt = m.patch_table(level, end_cap="gregory")
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    face, s, t = locations(m, k, level)
    want = m.limit_stencil_table(face, s, t, nw >= 3, nw >= 6, patch_table=pt)
    coords = m.find_patches(pt, face, s, t)
    dpt = osd.B200PatchTable.Create(pt)
    dst = osd.B200StencilTable.Create(st)
    lim = osd.B200StencilTable.CreateLimitStencils(dpt, dst, len(coords), coords_dev(coords), nw)
    return m, st, want, lim, coords


@pytest.mark.parametrize("nw", [6, 3, 1])
@pytest.mark.parametrize("shape,level", [("catmark_cube_creases0", 3), ("catmark_car", 2), ("catmark_gregory_test2", 3),
                                         ("catmark_nonquads", 3), ("catmark_pole64", 2), ("catmark_hole_test2", 3),
                                         ("catmark_edgecorner", 4), ("catmark_single_crease", 3)])
def test_catmark_tables_bit_identical_to_far(shape, level, nw):
    m, st, want, lim, coords = build(shape, level, nw)
    sizes, offsets, indices, ws = lim.ToHost(nw)
    assert lim.GetNumStencils() == want.num_stencils == int((coords["arrayIndex"] >= 0).sum())
    assert np.array_equal(sizes, want.sizes) and np.array_equal(offsets, want.offsets)
    assert np.array_equal(indices, want.indices)
    for k in range(nw):
        assert np.array_equal(ws[k].view(np.int32), getattr(want, STREAMS[k]).view(np.int32)), (shape, STREAMS[k])
    # the device-built table evaluates like Far's (bucketed layout, fused derivative streams)
    n = want.num_stencils
    src = dev(m.positions)
    far_tbl = osd.B200StencilTable.Create(want)
    a = torch.zeros((n, 3 * nw), device="cuda")
    b = torch.zeros((n, 3 * nw), device="cuda")
    aa, bb = [], []
    for k in range(nw):
        aa += [a, D(3 * k, 3, 3 * nw)]
        bb += [b, D(3 * k, 3, 3 * nw)]
    assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), *aa, lim)
    assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), *bb, far_tbl)
    assert torch.equal(a, b)


@pytest.mark.parametrize("shape,level", [("loop_icosahedron", 3), ("loop_cube_creases0", 2)])
def test_loop_tables_same_structure_close_weights(shape, level):
    """Triangle patches use the evaluation kernel's box-spline / Gregory-triangle forms (equivalent polynomials, another
    evaluation order than the reference): identical structure, weights equal to rounding."""
    m, st, want, lim, coords = build(shape, level, 6)
    sizes, offsets, indices, ws = lim.ToHost(6)
    assert np.array_equal(sizes, want.sizes) and np.array_equal(indices, want.indices)
    for k in range(6):
        w = getattr(want, STREAMS[k])
        wmax = np.maximum.reduceat(np.abs(w), offsets)
        worst = (np.maximum.reduceat(np.abs(ws[k] - w), offsets) / np.maximum(wmax, 1e-30)).max()
        assert worst <= (5e-6 if k == 0 else 2e-4), (shape, STREAMS[k], worst)


def test_unresolved_locations_and_reference_layout_only():
    """Holes produce no row (FindPatch == NULL); flags bit 0 keeps the table on the device (no read-back, CSR kernels)."""
    if not ref.available():
        pytest.skip("oracle/_ref/libosdref.so not present")
    m = ref.Mesh.from_shape("catmark_hole_test2")
    pt = m.patch_table(3, end_cap="gregory")
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    face, s, t = locations(m, 3000, 9)
    want = m.limit_stencil_table(face, s, t, True, True, patch_table=pt)
    coords = m.find_patches(pt, face, s, t)
    assert (coords["arrayIndex"] < 0).any()
    dpt = osd.B200PatchTable.Create(pt)
    dst = osd.B200StencilTable.Create(st)
    lim = osd.B200StencilTable.CreateLimitStencils(dpt, dst, len(coords), coords_dev(coords), 6, bucketed=False)
    assert lim.GetNumStencils() == want.num_stencils < len(coords)
    sizes, offsets, indices, ws = lim.ToHost(6)
    assert np.array_equal(indices, want.indices)
    n = want.num_stencils
    src = dev(m.positions)
    out = torch.zeros((n, 3), device="cuda")
    assert osd.B200Evaluator.EvalStencils(src, D(0, 3, 3), out, D(0, 3, 3), lim)
    exp = np.zeros((n, 3), np.float32)
    assert oracle.eval_stencils(m.positions.reshape(-1), (0, 3, 3), [exp.reshape(-1)], [(0, 3, 3)], want.sizes, want.offsets,
                                want.indices, [want.weights])
    assert np.abs(out.cpu().numpy() - exp).max() <= 1e-6 * max(1.0, np.abs(exp).max())
