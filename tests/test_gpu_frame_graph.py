"""the whole per-frame pipeline -- EvalStencils (refine + local points) -> FindPatches -> EvalPatches --
captured into ONE CUDA graph through the C ABI's stream arguments and replayed per frame; results must be bit-identical
to the eager calls.  (Every entry point is stream-ordered and allocation-free after its first call, so it can be captured.)"""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from tests.gpu_util import D, dev
from tests.util import golden, table_from, triple_from

pytestmark = pytest.mark.gpu


class _PT:
    def __init__(self, vertex, varying=None):
        self.vertex, self.varying, self.fvar = vertex, varying, []


@pytest.mark.parametrize("name", ["patches_catmark_car", "patches_catmark_gregory_test2", "patches_loop_icosahedron"])
def test_frame_graph_replay_matches_eager(name):
    d = golden(name)
    st = table_from(d, "st_")
    vtx = triple_from(d, "vtx_")
    var = triple_from(d, "var_") if "var_arrays" in d.files else None
    ncv, nst = st.num_control_verts, st.num_stencils
    stbl = osd.B200StencilTable.Create(st)
    pt = osd.B200PatchTable.Create(_PT(vtx, var))
    pm = osd.B200PatchMap.Create(_PT(vtx, var))
    coords = d["coords"]
    n = len(coords)
    face = dev((vtx.params["field0"][coords["patchIndex"]] & 0x0fffffff).astype(np.int32))
    s, t = dev(coords["s"]), dev(coords["t"])

    vb = torch.zeros((ncv + nst, 3), device="cuda")
    pc = torch.zeros(n * 5, dtype=torch.int32, device="cuda")
    found = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.zeros((n, 18), device="cuda")
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]

    def frame(stream):
        assert osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), stbl, deviceContext=stream)
        assert pm.FindPatches(n, face, s, t, pc, found, deviceContext=stream)
        assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, None, deviceContext=stream)

    def control_points(f):
        p = d["src0"].astype(np.float32).copy()
        ang = p[:, 2] * np.float32(np.sin(0.1 * f))
        c, sn = np.cos(ang), np.sin(ang)
        p[:, 0], p[:, 1] = p[:, 0] * c - p[:, 1] * sn, p[:, 0] * sn + p[:, 1] * c
        return dev(p.astype(np.float32))

    side = torch.cuda.Stream()
    vb[:ncv] = control_points(0)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        frame(side)                      # warm-up: first-call allocations (hull cache, basis tables) happen here
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        frame(side)
    for f in (1, 2, 17):
        cp = control_points(f)
        vb[:ncv] = cp
        out.fill_(-1.0)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        got = out.clone()
        got_vb = vb.clone()
        vb[ncv:] = 0
        out.fill_(-2.0)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            frame(side)
        side.synchronize()
        assert int(found.item()) == n
        assert torch.equal(got_vb, vb), f"{name} frame {f}: refined buffer differs between graph replay and eager"
        assert torch.equal(got, out), f"{name} frame {f}: limit outputs differ between graph replay and eager"
