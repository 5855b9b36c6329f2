"""Config-4 frame (60x catmark_car, adaptive L3, Gregory end caps, 10 M random samples) per patch variant:
   python tools/profile_config4.py            (needs oracle/_ref/libosdref.so for the Far tables)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import opensubdiv_b200 as osd  # noqa: E402
from opensubdiv_b200 import capi  # noqa: E402
from oracle import ref  # noqa: E402

D = osd.BufferDescriptor
n = int(os.environ.get("N", 10_000_000))
variants = [int(v) for v in os.environ.get("VARIANTS", "0,1,2,3").split(",")]   # grouping: 0 auto, 1 never, 2 always
m = ref.Mesh.from_shape_tiled("catmark_car", 60)
ptab = m.patch_table(3, end_cap="gregory", fvar=False, inf_sharp=True, legacy_sharp_corner=False)
st = m.stencil_table(intermediate_levels=True, patch_table=ptab)
ncv, nst = st.num_control_verts, st.num_stencils
vb = osd.B200VertexBuffer.Create(3, ncv + nst)
vb.UpdateData(np.ascontiguousarray(m.positions), 0, ncv)
stbl = osd.B200StencilTable.Create(st)
pt = osd.B200PatchTable.Create(ptab)
pm = osd.B200PatchMap.Create(ptab)
rng = np.random.default_rng(2024)
face = torch.from_numpy(rng.integers(0, m.num_ptex_faces, n).astype(np.int32)).cuda()
s = torch.from_numpy(rng.random(n, dtype=np.float32)).cuda()
t = torch.from_numpy(rng.random(n, dtype=np.float32)).cuda()
pc = torch.zeros(n * 5, dtype=torch.int32, device="cuda")
found = torch.zeros(1, dtype=torch.int32, device="cuda")
assert pm.FindPatches(n, face, s, t, pc)
out = torch.empty((n, 18), device="cuda")
args = []
for k in range(6):
    args += [out, D(3 * k, 3, 18)]
assert osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), stbl)


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(json.dumps({"patches": len(ptab.vertex.params), "stencils": nst, "coords": n,
                  "refine_ms": round(timed(lambda: osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), stbl)), 4),
                  "find_ms": round(timed(lambda: pm.FindPatches(n, face, s, t, pc)), 4),
                  "find_with_count_ms": round(timed(lambda: pm.FindPatches(n, face, s, t, pc, found)), 4)}), flush=True)
for order in ("random", "sorted"):
    if order == "sorted":
        rec = pc.view(n, 5)
        rec = rec[torch.argsort(rec[:, 1].to(torch.int64))].contiguous()
        pc = rec.view(-1)
    for v in variants:
        pt.SetVariant(v)
        ms = timed(lambda: osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, None))
        print(json.dumps({"order": order, "variant": v, "eval_patches_ms": round(ms, 4), "Gpts_per_s": round(n / ms / 1e6, 2),
                          "frac_of_hbm_920MB": round(n * 92 / (ms * 1e-3) / 1e9 / 6529.1, 3)}), flush=True)
    pt.SetVariant(0)
    # cached grouping (evaluator instance): the sort is paid once per coordinate set
    inst = osd.B200Evaluator.Create(D(0, 3, 3), D(0, 3, 18))
    bind_ms = timed(lambda: inst.BindPatchCoords(n, pc, pt), 5)
    ms = timed(lambda: osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, inst))
    print(json.dumps({"order": order, "variant": "cached plan", "bind_ms": round(bind_ms, 4), "eval_patches_ms": round(ms, 4),
                      "Gpts_per_s": round(n / ms / 1e6, 2), "frac_of_hbm_920MB": round(n * 92 / (ms * 1e-3) / 1e9 / 6529.1, 3)}), flush=True)
    del inst
