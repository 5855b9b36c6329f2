"""Standalone check of the C-ABI frame capture (b200osd_frame_*): run as a subprocess by tests/test_gpu_zz_frame_capi.py
so that a failure cannot disturb the rest of the GPU suite.  Records EvalStencils -> FindPatches -> EvalPatches on the
frame's own stream, replays it for new control points and compares bit for bit with eager evaluation."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import opensubdiv_b200 as osd  # noqa: E402
from tests.util import golden, table_from, triple_from  # noqa: E402


class _PT:
    def __init__(self, vertex, varying=None):
        self.vertex, self.varying, self.fvar = vertex, varying, []


def main(name="patches_catmark_car"):
    D = osd.BufferDescriptor
    d = golden(name)
    st, vtx = table_from(d, "st_"), triple_from(d, "vtx_")
    var = triple_from(d, "var_") if "var_arrays" in d.files else None
    ncv, nst = st.num_control_verts, st.num_stencils
    stbl = osd.B200StencilTable.Create(st)
    pt = osd.B200PatchTable.Create(_PT(vtx, var))
    pm = osd.B200PatchMap.Create(_PT(vtx, var))
    coords = d["coords"]
    n = len(coords)
    face = torch.from_numpy((vtx.params["field0"][coords["patchIndex"]] & 0x0fffffff).astype(np.int32)).cuda()
    s, t = torch.from_numpy(coords["s"].copy()).cuda(), torch.from_numpy(coords["t"].copy()).cuda()
    vb = torch.zeros((ncv + nst, 3), device="cuda")
    pc = torch.zeros(n * 5, dtype=torch.int32, device="cuda")
    out = torch.zeros((n, 18), device="cuda")
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]

    def run(ctx):
        assert osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), stbl, deviceContext=ctx)
        assert pm.FindPatches(n, face, s, t, pc, None, deviceContext=ctx)
        assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, None, deviceContext=ctx)

    def control_points(f):
        p = d["src0"].astype(np.float32).copy()
        p[:, 0] += np.float32(0.01 * f) * p[:, 2]
        return torch.from_numpy(p).cuda()

    frame = osd.B200FrameGraph.Create()
    assert frame is not None and frame.cuda_stream
    vb[:ncv] = control_points(0)
    torch.cuda.synchronize()
    run(frame)                                   # eager warm-up on the frame's stream
    assert frame.Synchronize()
    assert frame.Begin()
    run(frame)
    assert frame.End()
    for f in (1, 2, 5):
        vb[:ncv] = control_points(f)
        vb[ncv:] = 0
        out.fill_(-1.0)
        torch.cuda.synchronize()
        assert frame.Launch()
        assert frame.Synchronize()
        got_vb, got = vb.clone(), out.clone()
        vb[ncv:] = 0
        out.fill_(-2.0)
        torch.cuda.synchronize()
        run(None)
        torch.cuda.synchronize()
        assert torch.equal(got_vb, vb), f"frame {f}: refined buffer differs"
        assert torch.equal(got, out), f"frame {f}: limit outputs differ"
    print("FRAME CAPI OK")


if __name__ == "__main__":
    main()
