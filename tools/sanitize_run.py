"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
     compute-sanitizer --tool racecheck python tools/sanitize_run.py
Golden fixtures only (seconds under the sanitizer); results are checked against the fixtures' reference outputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import opensubdiv_b200 as osd  # noqa: E402
from tests.util import golden, table_from, triple_from, assert_close  # noqa: E402
from tests.gpu_util import oracle_patches, oracle_stencils, coords_dev, dev  # noqa: E402

D = osd.BufferDescriptor
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")


class _PT:
    def __init__(self, vertex, varying=None, fvar=None):
        self.vertex, self.varying, self.fvar = vertex, varying, fvar or []


def stencils():
    d = golden("stencils_catmark_car")
    t = table_from(d, "t_")
    ncv, n = t.num_control_verts, t.num_stencils
    scale = oracle_stencils(d["src"], (0, 3, 3), n, 3, t, 1, abs_scale=True)[0]
    tbl = osd.B200StencilTable.Create(t)
    for v in (0, 1, 8, 113, 122):                      # one-shot, CSR, persistent, TMA-staged rings
        tbl.SetVariant(v)
        vb = osd.B200VertexBuffer.Create(3, ncv + n)   # Osd::Mesh::Refine layout: src and dst in one buffer
        vb.UpdateData(np.ascontiguousarray(d["src"], np.float32), 0, ncv)
        assert osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), tbl)
        osd.B200Evaluator.Synchronize()
        assert_close(vb.as_tensor()[ncv:].cpu().numpy(), d["out"], scale, f"stencils variant {v}")
    d = golden("limit_catmark_torus")
    t = table_from(d, "t_")
    n = t.num_stencils
    tbl = osd.B200StencilTable.Create(t)
    scales = oracle_stencils(d["src"], (0, 3, 3), n, 3, t, 6, abs_scale=True)
    for v in (0, 8, 121, 122):                         # K = 6: TMA default, persistent, other ring shapes
        tbl.SetVariant(v)
        out = torch.zeros((n, 18), device="cuda")
        a = []
        for k in range(6):
            a += [out, D(3 * k, 3, 18)]
        assert osd.B200Evaluator.EvalStencils(dev(d["src"]), D(0, 3, 3), *a, tbl)
        res = out.cpu().numpy()
        for k in range(6):
            assert_close(res[:, 3 * k:3 * k + 3], d["out_" + OUT6[k]], scales[k], f"limit stencils variant {v} {OUT6[k]}")
    # batched instances
    d = golden("stencils_catmark_car")
    t = table_from(d, "t_")
    ncv, n = t.num_control_verts, t.num_stencils
    tbl = osd.B200StencilTable.Create(t)
    pitch = (ncv + n) * 3
    host = np.zeros((5, ncv + n, 3), np.float32)
    host[:, :ncv] = d["src"]
    b = dev(host.reshape(-1))
    assert osd.B200Evaluator.EvalStencilsBatched(b, D(0, 3, 3), b, D(ncv * 3, 3, 3), tbl, 5, pitch)
    torch.cuda.synchronize()


def patches():
    for name in ("patches_catmark_car", "patches_loop_icosahedron", "patches_catmark_fvar_bound1"):
        d = golden(name)
        vtx = triple_from(d, "vtx_")
        var = triple_from(d, "var_") if "var_arrays" in d.files else None
        fv = triple_from(d, "fvar_") if "fvar_arrays" in d.files else None
        pt = osd.B200PatchTable.Create(_PT(vtx, var, [fv] if fv is not None else []))
        coords = d["coords"]
        n = len(coords)
        pc = coords_dev(coords)
        src = dev(d["vb"])
        scales = oracle_patches(d["vb"], (0, 3, 3), 3, coords, vtx, 6, abs_scale=True)
        for variant in (1, 2, 3):                      # caller's order (staged hulls), grouped per call, hull cache
            pt.SetVariant(variant)
            out = torch.zeros((n, 18), device="cuda")
            a = []
            for k in range(6):
                a += [out, D(3 * k, 3, 18)]
            assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *a, n, pc, pt, None)
            res = out.cpu().numpy()
            for k in range(6):
                assert_close(res[:, 3 * k:3 * k + 3], d["out_" + OUT6[k]], scales[k], f"{name} variant {variant} {OUT6[k]}")
        pt.SetVariant(0)
        if name == "patches_catmark_car":              # the general kernel instantiation with the true Gregory derivatives
            g = golden("truederiv_catmark_car")
            pt.SetGregoryTrueDerivatives(True)
            for variant in (1, 3):
                pt.SetVariant(variant)
                out = torch.zeros((n, 18), device="cuda")
                a = []
                for k in range(6):
                    a += [out, D(3 * k, 3, 18)]
                assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *a, n, pc, pt, None)
                assert np.abs(out.cpu().numpy()[:, 3:6] - g["out_du"]).max() <= 1e-3 * max(1.0, np.abs(g["out_du"]).max())
            pt.SetVariant(0)
            pt.SetGregoryTrueDerivatives(False)
        inst = osd.B200Evaluator.Create(D(0, 3, 3), D(0, 3, 18))
        assert inst.BindPatchCoords(n, pc, pt)
        out = torch.zeros((n, 3), device="cuda")
        assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), out, D(0, 3, 3), n, pc, pt, inst)
        if var is not None:
            o = torch.zeros((n, 3), device="cuda")
            assert osd.B200Evaluator.EvalPatchesVarying(dev(d["var_vb"]), D(0, 3, 3), o, D(0, 3, 3), n, pc, pt, None)
        if fv is not None:
            o = torch.zeros((n, 2), device="cuda")
            assert osd.B200Evaluator.EvalPatchesFaceVarying(dev(d["fvar_vb"]), D(0, 2, 2), o, D(0, 2, 2), n, pc, pt, 0, None)
        torch.cuda.synchronize()


def patch_map_and_limit_builder():
    g = golden("patchmap_catmark_car")
    pm = osd.B200PatchMap.Create(_PT(type("T", (), dict(arrays=g["arrays"], params=g["params"]))()),
                                 patchesAreTriangular=bool(g["triangular"]))
    ns = len(g["face"])
    rec = torch.zeros(ns * 5, dtype=torch.int32, device="cuda")
    found = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert pm.FindPatches(ns, dev(g["face"]), dev(g["s"]), dev(g["t"]), rec, found)
    torch.cuda.synchronize()
    d = golden("patches_catmark_car")
    st, vtx = table_from(d, "st_"), triple_from(d, "vtx_")
    pt = osd.B200PatchTable.Create(_PT(vtx))
    cv = osd.B200StencilTable.Create(st)
    coords = d["coords"]
    lim = osd.B200StencilTable.CreateLimitStencils(pt, cv, len(coords), coords_dev(coords), 6)
    out = torch.zeros((lim.GetNumStencils(), 18), device="cuda")
    a = []
    for k in range(6):
        a += [out, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalStencils(dev(d["src0"]), D(0, 3, 3), *a, lim)
    torch.cuda.synchronize()
    ref = d["out_p"]
    assert np.abs(out[:, 0:3].cpu().numpy() - ref).max() <= 2e-5 * max(1.0, float(np.abs(ref).max()))


if __name__ == "__main__":
    stencils()
    patches()
    patch_map_and_limit_builder()
    print("SANITIZE RUN OK")
