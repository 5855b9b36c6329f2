"""EvalPatches microbenchmark on the synthetic torus (10 M coords, 100 k regular patches, xyz):
     python tools/profile_patches.py            # timings: {random, sorted} x {interleaved record, separate buffers} x nOut
     NCU=1 ncu --set full -k regex:patch_kernel ... python tools/profile_patches.py    # one launch per order, 6 outputs"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import opensubdiv_b200 as osd  # noqa: E402
from opensubdiv_b200 import capi, synth  # noqa: E402

D = osd.BufferDescriptor
n = int(os.environ.get("N", 10_000_000))
ncu = os.environ.get("NCU") == "1"
variants = [int(v) for v in os.environ.get("VARIANTS", "0").split(",")]
mesh = synth.torus_quads(400, 250)
ptab = synth.torus_patch_table(mesh)
pt = osd.B200PatchTable.Create(ptab)
if ncu:
    pt.SetVariant(int(os.environ.get("VARIANTS", "0").split(",")[0]))
src = torch.from_numpy(np.ascontiguousarray(synth.deform(mesh.positions, 1))).cuda()


def outputs(nw, interleaved):
    if interleaved:
        out = torch.empty((n, 3 * nw), device="cuda")
        a = []
        for k in range(nw):
            a += [out, D(3 * k, 3, 3 * nw)]
        return a, out
    outs = [torch.empty((n, 3), device="cuda") for _ in range(nw)]
    a = []
    for o in outs:
        a += [o, D(0, 3, 3)]
    return a, outs


for sort in (False, True):
    coords = synth.random_patch_coords(len(mesh.faces), n, seed=2024, sort_by_patch=sort)
    pc = torch.from_numpy(coords.view(np.uint8)).cuda()
    if ncu:
        a, keep = outputs(6, True)
        assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *a, n, pc, pt, None)
        torch.cuda.synchronize()
        continue
    for variant in variants:
        pt.SetVariant(variant)
        for nw in (1, 3, 6):
            for inter in (True, False):
                if nw == 1 and not inter:
                    continue
                a, keep = outputs(nw, inter)
                for _ in range(3):
                    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *a, n, pc, pt, None)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *a, n, pc, pt, None)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 20
                print(json.dumps({"order": "sorted" if sort else "random", "variant": variant, "nOut": nw,
                                  "layout": "interleaved" if inter else "separate", "ms": round(ms, 4),
                                  "Gpts_per_s": round(n / ms / 1e6, 2)}), flush=True)
                del a, keep
    pt.SetVariant(0)
print("done")
