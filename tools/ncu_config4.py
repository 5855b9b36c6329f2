"""One config-4 EvalPatches call per mode for ncu (random coords: grouped per call; sorted coords: caller order).
   ncu --set full -k regex:'patch_run|bin_' -o gpurun_out/x python tools/ncu_config4.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import opensubdiv_b200 as osd  # noqa: E402
from oracle import ref  # noqa: E402

D = osd.BufferDescriptor
n = int(os.environ.get("N", 10_000_000))
m = ref.Mesh.from_shape_tiled("catmark_car", int(os.environ.get("TILES", 60)))
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
assert pm.FindPatches(n, face, s, t, pc)
out = torch.empty((n, 18), device="cuda")
args = []
for k in range(6):
    args += [out, D(3 * k, 3, 18)]
assert osd.B200Evaluator.EvalStencils(vb, D(0, 3, 3), vb, D(ncv * 3, 3, 3), stbl)
mode = os.environ.get("MODE", "auto")
if mode == "all":
    pt.SetVariant(2)
    assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, None)   # random, grouped per call
    pt.SetVariant(1)
    assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, None)   # random, caller order
pt.SetVariant(0)
assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pc, pt, None)       # random, automatic (probe -> hull cache)
rec = pc.view(n, 5)
pcs = rec[torch.argsort(rec[:, 1].to(torch.int64))].contiguous().view(-1)
assert osd.B200Evaluator.EvalPatches(vb, D(0, 3, 3), *args, n, pcs, pt, None)      # sorted, automatic (probe -> caller order)
torch.cuda.synchronize()
