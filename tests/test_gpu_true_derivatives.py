"""B200OSD_PATCH_GREGORY_TRUE_DERIVATIVES on the GPU (the reference's build option OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES,
osd/patchBasis.h:421-487) against golden outputs of the reference compiled with it (tests/golden/truederiv_*.npz):
EvalPatches through the table in every serving mode, the raw entry with an option mask, a patch plan, and the
device-built limit-stencil table (bit-identical to Far::LimitStencilTableFactory of the switched build)."""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from oracle import oracle, ref
from tests.gpu_util import D, dev, coords_dev
from tests.util import golden, triple_from, assert_close, assert_close_bbox

pytestmark = pytest.mark.gpu
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")
STREAMS = ("weights", "du", "dv", "duu", "duv", "dvv")


class _PT:
    def __init__(self, vertex):
        self.vertex, self.varying, self.fvar = vertex, None, []


def switched_scales(d, tr, nw):
    coords, vb = d["coords"], d["vb"]
    outs = [np.zeros((len(coords), 3), np.float32) for _ in range(nw)]
    with oracle.gregory_true_derivatives(), oracle.abs_mode(2):
        assert oracle.eval_patches(np.ascontiguousarray(vb).reshape(-1), (0, 3, 3), [o.reshape(-1) for o in outs],
                                   [(0, 3, 3)] * nw, coords, tr.arrays, tr.indices, tr.params)
    return outs


@pytest.mark.parametrize("shape", ["catmark_cube_creases0", "catmark_gregory_test2", "catmark_car"])
def test_eval_patches_with_true_gregory_derivatives(shape):
    d, g = golden("patches_" + shape), golden("truederiv_" + shape)
    vtx = triple_from(d, "vtx_")
    coords = d["coords"]
    n = len(coords)
    pc, src = coords_dev(coords), dev(d["vb"])
    pt = osd.B200PatchTable.Create(_PT(vtx))
    assert not pt.GetGregoryTrueDerivatives()
    pt.SetGregoryTrueDerivatives(True)
    assert pt.GetGregoryTrueDerivatives()
    scales = switched_scales(d, vtx, 6)
    bbox = float((d["vb"].max(axis=0) - d["vb"].min(axis=0)).max())
    depth = (vtx.params["field1"][coords["patchIndex"]] & 0xF).astype(np.int64)
    order_of = (0, 1, 1, 2, 2, 2)
    first = {}
    for variant in (0, 1, 2, 3):
        for nw in (1, 3, 6):
            out = torch.full((n, 3 * nw), float("nan"), device="cuda")
            args = []
            for k in range(nw):
                args += [out, D(3 * k, 3, 3 * nw)]
            pt.SetVariant(variant)
            assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, None)
            res = out.cpu().numpy()
            for k in range(nw):
                assert_close(res[:, 3 * k:3 * k + 3], g["out_" + OUT6[k]], scales[k], f"{shape} variant={variant} nw={nw} {OUT6[k]}")
                # scale-free gate; the true second derivatives carry 1/(s+t)^2 factors next to patch corners, so the bound
                # of the default path (2.5e-6) is kept for them and 1e-6 for P and D1
                assert_close_bbox(res[:, 3 * k:3 * k + 3], g["out_" + OUT6[k]], bbox, depth, order_of[k],
                                  f"{shape} variant={variant} nw={nw} {OUT6[k]} (bbox gate)", tol=1e-6 if order_of[k] < 2 else 2.5e-6)
            if nw == 6:
                first.setdefault("ref", res)
                assert np.array_equal(res, first["ref"]), f"variant {variant} differs bitwise from variant 0"
    pt.SetVariant(0)
    # the option changes derivatives only, and by a lot on these shapes
    assert np.abs(first["ref"][:, 3:6] - d["out_du"]).max() > 1e-3
    # raw entry with the option mask; without it the default approximation comes out
    outs = [torch.zeros((n, 3), device="cuda") for _ in range(6)]
    raw = (n, pc, pt.GetPatchArrayBuffer(), pt.GetPatchIndexBuffer(), pt.GetPatchParamBuffer())
    assert osd.B200Evaluator.EvalPatchesRaw(src, D(0, 3, 3), [(o, D(0, 3, 3)) for o in outs], *raw, gregory_true_derivatives=True)
    for k in range(6):
        assert np.array_equal(outs[k].cpu().numpy(), first["ref"][:, 3 * k:3 * k + 3]), f"raw {OUT6[k]}"
    assert osd.B200Evaluator.EvalPatchesRaw(src, D(0, 3, 3), [(o, D(0, 3, 3)) for o in outs], *raw)
    for k in range(6):
        assert np.abs(outs[k].cpu().numpy() - d["out_" + OUT6[k]]).max() <= 1e-3 * max(1.0, np.abs(d["out_" + OUT6[k]]).max())
    # a bound evaluator instance (patch plan) honours the table's option
    inst = osd.B200Evaluator.Create(D(0, 3, 3), D(0, 3, 18), D(3, 3, 18), D(6, 3, 18), D(9, 3, 18), D(12, 3, 18), D(15, 3, 18))
    out = torch.zeros((n, 18), device="cuda")
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]
    assert inst.BindPatchCoords(n, pc, pt)
    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, inst)
    assert np.array_equal(out.cpu().numpy(), first["ref"])


@pytest.mark.parametrize("shape", ["catmark_gregory_test2", "catmark_cube_creases0"])
def test_device_limit_table_with_true_gregory_derivatives(shape):
    if not ref.available():
        pytest.skip("oracle/_ref/libosdref.so not present")
    d, g = golden("limit_" + shape), golden("truederiv_limit_" + shape)
    m = ref.Mesh.from_shape(shape).refine_adaptive(3)
    pt = m.patch_table(3, end_cap="gregory", refine_first=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    coords = m.find_patches(pt, d["face"], d["s"], d["t"])
    dpt = osd.B200PatchTable.Create(pt)
    dpt.SetGregoryTrueDerivatives(True)
    dst = osd.B200StencilTable.Create(st)
    lim = osd.B200StencilTable.CreateLimitStencils(dpt, dst, len(coords), coords_dev(coords), 6)
    sizes, offsets, indices, ws = lim.ToHost(6)
    assert np.array_equal(sizes, g["t_sizes"]) and np.array_equal(indices, g["t_indices"])
    for k in range(6):
        assert np.array_equal(ws[k].view(np.int32), g["t_" + STREAMS[k]].view(np.int32)), (shape, STREAMS[k])
    # without the option the default table comes out (bit-identical to the plain fixture)
    dpt.SetGregoryTrueDerivatives(False)
    lim = osd.B200StencilTable.CreateLimitStencils(dpt, dst, len(coords), coords_dev(coords), 6)
    ws = lim.ToHost(6)[3]
    for k in range(6):
        assert np.array_equal(ws[k].view(np.int32), d["t_" + STREAMS[k]].view(np.int32)), (shape, STREAMS[k])


def test_unknown_option_bits_are_rejected():
    d = golden("patches_catmark_cube_creases0")
    pt = osd.B200PatchTable.Create(_PT(triple_from(d, "vtx_")))
    from opensubdiv_b200 import capi
    assert capi.lib().b200osd_patch_table_set_options(pt._h, 2) == capi.ERR_INVALID
    assert capi.lib().b200osd_patch_table_get_options(pt._h) == 0
