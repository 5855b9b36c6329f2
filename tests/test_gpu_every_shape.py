"""Every shape of the reference's regression list (regression/osd_regression/main.cpp:250-330 walks the same list) through
the GPU path, against the unmodified reference on the box (oracle/_ref/libosdref.so travels with the snapshot):
Osd::Mesh-style refinement in one buffer (fast path within the rounding of a row's length, reference-exact mode
bit-identical to Osd::CpuEvaluator), device FindPatches bit-exact, EvalPatches P + D1 + D2 in every serving mode.
Skipped where the reference build is absent."""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from oracle import oracle, ref
from tests.gpu_util import D, coords_dev
from tests.util import assert_close

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libosdref.so not present")]
SHAPES = ref.shape_names() if ref.available() else []
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")


@pytest.mark.parametrize("shape", SHAPES)
def test_shape_refine_locate_evaluate(shape):
    m = ref.Mesh.from_shape(shape)
    bilinear = shape.startswith("bilinear")
    level = 1 if shape.endswith("pole360") else 3
    pt = m.patch_table(level, end_cap="gregory", inf_sharp=True, legacy_sharp_corner=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    ncv, n = st.num_control_verts, st.num_stencils
    # reference: refine in one buffer (osd/mesh.h:505-519)
    vb = np.zeros((ncv + n, 3), np.float32)
    vb[:ncv] = m.positions
    assert ref.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st)
    scale = np.zeros((n, 3), np.float32)
    with oracle.abs_mode(1):
        oracle.eval_stencils(vb.reshape(-1), (0, 3, 3), [scale.reshape(-1)], [(0, 3, 3)], st.sizes, st.offsets, st.indices, [st.weights])
    # GPU: the same through the table handle
    gvb = osd.B200VertexBuffer.Create(3, ncv + n)
    gvb.UpdateData(np.ascontiguousarray(m.positions), 0, ncv)
    tbl = osd.B200StencilTable.Create(st)
    assert osd.B200Evaluator.EvalStencils(gvb, D(0, 3, 3), gvb, D(ncv * 3, 3, 3), tbl)
    got = gvb.as_tensor().cpu().numpy()
    # Fast path: fused multiply-adds, short rows summed in control-index order.  Two correct fp32 summations of n terms
    # differ by up to ~n * 2^-23 of sum|w||x| (catmark_bishop has a 25-term row where the reference is +5.9e-7 and the FMA
    # chain -5.8e-7 away from the exact sum): 1e-6 for rows of up to 8 terms, n * 2^-23 beyond.
    tol = np.maximum(1e-6, st.sizes.astype(np.float64) * 2.0 ** -23)[:, None]
    err = np.abs(got[ncv:].astype(np.float64) - vb[ncv:]) / np.maximum(np.maximum(np.abs(vb[ncv:]), scale), 1e-30)
    assert (err <= tol).all(), f"{shape} refine: worst error / tolerance {(err / tol).max():.3f}"
    # Reference-exact mode: rounded product then add, table order -- bit-identical to Osd::CpuEvaluator for every row
    exact = osd.B200StencilTable.Create(st, reference_exact=True)
    gvb.UpdateData(np.zeros((n, 3), np.float32), ncv, n)
    assert osd.B200Evaluator.EvalStencils(gvb, D(0, 3, 3), gvb, D(ncv * 3, 3, 3), exact)
    got = gvb.as_tensor().cpu().numpy()
    assert np.array_equal(got[ncv:].view(np.int32), vb[ncv:].view(np.int32)), f"{shape}: reference-exact mode is not bit-identical"
    if bilinear:
        return
    gvb.UpdateData(np.ascontiguousarray(vb[ncv:]), ncv, n)       # the patch checks run on the reference-refined buffer
    rng = np.random.default_rng(17)
    k = 2000
    face = rng.integers(0, m.num_ptex_faces, k).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    s[::37], t[::41] = 0.0, 1.0
    if m.reg_face_size == 3:
        flip = s + t >= 1
        s, t = np.where(flip, 1 - s, s).astype(np.float32), np.where(flip, 1 - t, t).astype(np.float32)
    want = m.find_patches(pt, face, s, t)
    dpt = osd.B200PatchTable.Create(pt)
    pm = osd.B200PatchMap.Create(pt)
    pc = torch.zeros(k * 5, dtype=torch.int32, device="cuda")
    assert pm.FindPatches(k, torch.from_numpy(face).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(t).cuda(), pc)
    found = pc.view(k, 5).cpu().numpy()
    hit = want["arrayIndex"] >= 0
    assert np.array_equal(found[:, 0] >= 0, hit), shape
    assert np.array_equal(found[hit], np.ascontiguousarray(want[hit]).view(np.int32).reshape(-1, 5)), f"{shape} FindPatches"
    live = np.ascontiguousarray(want[hit])
    x = [np.zeros((len(live), 3), np.float32) for _ in range(6)]
    z = [np.zeros((len(live), 3), np.float32) for _ in range(6)]
    assert ref.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in x], [(0, 3, 3)] * 6, live, pt.vertex)
    with oracle.abs_mode(2):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in z], [(0, 3, 3)] * 6, live, pt.vertex.arrays,
                            pt.vertex.indices, pt.vertex.params)
    # evaluate at the GPU-located coordinates (holes keep their NaN)
    first = None
    for variant in (1, 3, 2):
        dpt.SetVariant(variant)
        out = torch.full((k, 18), float("nan"), device="cuda")
        args = []
        for q in range(6):
            args += [out, D(3 * q, 3, 18)]
        assert osd.B200Evaluator.EvalPatches(gvb, D(0, 3, 3), *args, k, pc, dpt, None)
        res = out.cpu().numpy()
        assert np.isnan(res[~hit]).all()
        if first is None:
            first = res
            for q in range(6):
                assert_close(res[hit, 3 * q:3 * q + 3], x[q], z[q], f"{shape} {OUT6[q]}")
        else:
            assert np.array_equal(res[hit], first[hit]), f"{shape}: variant {variant} differs bitwise from the caller's order"
