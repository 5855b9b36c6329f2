#!/usr/bin/env python
"""Device-time sweep of every kernel variant on the BASELINE configs (run under gpurun; not a bench line).

    python tools/gpu_sweep.py [--quick] [--only stencil|patch] [--iters N]

Prints one JSON object per measurement to stdout (and gpurun_out/sweep.jsonl).  All timings are CUDA events around
`iters` back-to-back calls after warm-up; inputs are far larger than L2 for the stencil tables.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import opensubdiv_b200 as osd          # noqa: E402
from opensubdiv_b200 import capi, synth  # noqa: E402

D = osd.BufferDescriptor
PEAK = 6529.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
OUT = None


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if OUT:
        OUT.write(line + "\n")
        OUT.flush()


def time_calls(fn, iters, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def stencil_case(name, table, Ls, nouts, iters, variants=(0, 1, 2, 8, 11, 12), locality=False, idx16=True, sort_elements=False):
    lib = capi.lib()
    t0 = time.time()
    tbl = osd.B200StencilTable.Create(table, locality=locality, idx16=idx16, sort_elements=sort_elements)
    assert tbl is not None, capi.last_error()
    build_s = time.time() - t0
    ncv, n = table.num_control_verts, table.num_stencils
    for L in Ls:
        for nout in nouts:
            src = torch.randn((ncv, L), device="cuda")
            out = torch.empty((n, L * nout), device="cuda")
            args = []
            for k in range(nout):
                args += [out, D(L * k, L, L * nout)]
            alg = table.algorithmic_bytes(nout, L, L)
            for v in variants:
                tbl.SetVariant(v)
                ms = time_calls(lambda: osd.B200Evaluator.EvalStencils(src, D(0, L, L), *args, tbl), iters)
                tbl.SetVariant(0)
                emit(case=name, kind="stencil", locality=locality, idx16=idx16, sorted_elems=sort_elements, L=L, nout=nout, variant=v, ms=ms, rows=n, elements=table.num_elements,
                     gverts_per_s=n / ms / 1e6, alg_MB=alg / 1e6, alg_GBps=alg / ms / 1e6, frac_of_measured_peak=alg / ms / 1e6 / PEAK,
                     stream_MB=tbl.GetStreamBytes(nout) / 1e6, table_build_s=build_s)
    del tbl
    torch.cuda.empty_cache()


def patch_case(name, mesh, n, iters):
    ptab = synth.torus_patch_table(mesh)
    pt = osd.B200PatchTable.Create(ptab)
    src = torch.from_numpy(mesh.positions).cuda()
    for order_name, sort in (("random", False), ("sorted_by_patch", True)):
        coords = synth.random_patch_coords(len(mesh.faces), n, seed=2024, sort_by_patch=sort)
        pc = torch.from_numpy(coords.view(np.uint8)).cuda()
        for nout in (1, 3, 6):
            out = torch.empty((n, 3 * nout), device="cuda")
            args = []
            for k in range(nout):
                args += [out, D(3 * k, 3, 3 * nout)]
            alg = n * (20 + nout * 12)
            for pv in (0, 1, 2, 3):
                pt.SetVariant(pv)
                ms = time_calls(lambda: osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, None), iters)
                pt.SetVariant(0)
                emit(case=name, kind="patch", path={0: "auto", 1: "caller_order", 2: "grouped_per_call", 3: "hull_cache"}[pv], coords=n, order=order_name,
                     nout=nout, ms=ms, gpts_per_s=n / ms / 1e6, alg_MB=alg / 1e6, alg_GBps=alg / ms / 1e6,
                     frac_of_measured_peak=alg / ms / 1e6 / PEAK)
        # face-varying-like: 2 floats through the linear (QUADS) varying patches, value only
        uv = torch.rand((mesh.num_verts, 2), device="cuda")
        o2 = torch.empty((n, 2), device="cuda")
        ms = time_calls(lambda: osd.B200Evaluator.EvalPatchesVarying(uv, D(0, 2, 2), o2, D(0, 2, 2), n, pc, pt, None), iters)
        emit(case=name, kind="patch_varying_uv", coords=n, order=order_name, nout=1, ms=ms, gpts_per_s=n / ms / 1e6,
             alg_MB=n * 28 / 1e6, alg_GBps=n * 28 / ms / 1e6, frac_of_measured_peak=n * 28 / ms / 1e6 / PEAK)


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--far", action="store_true", help="config 2 / 5 tables from the reference's Far::StencilTableFactory (its row order)")
    ap.add_argument("--variants", default="0,122")
    ap.add_argument("--sort", action="store_true", help="sort each row's elements by control index at table creation")
    a = ap.parse_args()
    variants = tuple(int(v) for v in a.variants.split(","))

    def uniform_table(mesh, level, scheme):
        if a.far:
            from oracle import ref as oref
            m = oref.Mesh.from_topology(scheme, mesh.num_verts, np.full(len(mesh.faces), mesh.faces.shape[1], np.int32),
                                        mesh.faces.reshape(-1))
            far = m.refine_uniform(level).stencil_table()
            return synth.SynthStencilTable(num_control_verts=far.num_control_verts, sizes=far.sizes, offsets=far.offsets,
                                           indices=far.indices, weights=far.weights)
        return synth.uniform_stencil_table(mesh, level)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    OUT = open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "a")
    emit(kind="env", gpu=torch.cuda.get_device_name(0), peak_GBps=PEAK, version=capi.lib().b200osd_version().decode())
    if a.only in ("", "stencil"):
        mesh = synth.torus_quads(400, 250)
        table = uniform_table(mesh, 3, "catmark")
        tag = "_far_order" if a.far else "_index_sorted"
        if a.quick:
            stencil_case("cfg2_catmark_400x250_L3" + tag, table, (6, 8), (1,), 20, variants=variants, sort_elements=a.sort)
        else:
            stencil_case("cfg2_catmark_400x250_L3" + tag, table, (6, 3, 4, 8), (1,), a.iters, variants=variants)

            del table
            rng = np.random.default_rng(12345)
            face = np.sort(rng.integers(0, len(mesh.faces), 1_000_000)).astype(np.int32)
            ls = synth.torus_limit_stencil_table(mesh, face, rng.random(1_000_000, dtype=np.float32),
                                                 rng.random(1_000_000, dtype=np.float32))
            stencil_case("cfg3_limit_1M_x16", ls, (3,), (1, 3, 6), a.iters, variants=variants)
            del ls
            mesh5 = synth.torus_tris(1000, 500)
            t5 = uniform_table(mesh5, 2, "loop")
            stencil_case("cfg5_loop_1000x500_L2" + tag, t5, (3,), (1,), a.iters, variants=variants)

            del t5
    if a.only in ("", "patch") and not a.quick:
        patch_case("cfg4_torus_regular_10M", synth.torus_quads(400, 250), 10_000_000, max(10, a.iters // 5))


if __name__ == "__main__":
    main()
