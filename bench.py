#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Osd evaluator (BASELINE.json metric: refined verts/s, EvalStencils).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic deforming Catmark quad torus 400x250 (100 000 control vertices),
uniform level 3, last level only = 6.4 M stencil rows / 84.1 M elements, 6-float interleaved xyz+normal.
One step = one frame = one EvalStencils pass over the whole table.

  value      refined verts/s, control points already resident in HBM, CUDA-event timed (max over ranks)
  e2e        the same metric through the C ABI with HOST buffers: per step UpdateData (pinned H2D of the control
             points) + EvalStencils + ReadData (D2H of the refined vertices) + Synchronize
  roofline   algorithmic bytes (reference table formats) / device time per step vs measured HBM peak
  cpu_baseline  the reference's own Osd::CpuEvaluator / OmpEvaluator (oracle/_ref, compiled in place from
             /root/reference) on this box's host cores, same table, bounded number of frames; falls back to the
             C oracle port when the reference build is absent
  eval_patches  (N = 1; N > 1 with --shard-patches) the second half of the metric: limit pts/s of EvalPatches with 1st +
             2nd derivatives on 10 M PatchCoords (random and patch-sorted), FindPatches on a real adaptive table, and the
             reference's CPU evaluators on a sample of the same coordinates
  incumbent_cuda  (N = 1) the reference's own CUDA kernels (osd/cudaKernel.cu compiled for sm_100a under oracle/_ref) on
             the same device buffers: a reported baseline like cpu_baseline

Multi-GPU (N > 1): weak scaling.  The scene is N such meshes; rank r owns the stencil rows of mesh r (row-range
sharding, tables pre-sharded, no collective on the table side); every frame rank 0's deformed control points of the
WHOLE scene (N x 2.4 MB) are replicated with one NCCL broadcast -- the only exchange step -- and then each rank
evaluates its rows.  value = all rows of all ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NU, NV, LEVEL, L = 400, 250, 3, 6
WORKLOAD = "catmark_torus_400x250_uniform_L3_laststencils_xyz+normal_f32x6"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------ clocks --
class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------- workload --
def build_workload():
    from opensubdiv_b200 import synth
    t0 = time.time()
    mesh = synth.torus_quads(NU, NV)
    table = synth.uniform_stencil_table(mesh, LEVEL)
    log(f"[bench] table built in {time.time() - t0:.1f}s: {table.num_stencils} rows, {table.num_elements} elements")
    return mesh, table


def frame_primvars(mesh, frame):
    from opensubdiv_b200 import synth
    p = synth.deform(mesh.positions, frame)
    return np.ascontiguousarray(np.concatenate([p, synth.vertex_normals_like(p)], axis=1), dtype=np.float32)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu summary, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("sell_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------- reference arm --
def cpu_reference_frames(mesh, table, frames: int, threads: int, prefer: str = "auto"):
    """Times the reference's own CPU evaluators on full frames of the workload.  Returns a dict with verts/s."""
    from oracle import ref as oref
    src = frame_primvars(mesh, 0)
    n = table.num_stencils
    dst = np.zeros((n, L), np.float32)
    res = {}
    if oref.available():
        from types import SimpleNamespace
        t = SimpleNamespace(sizes=table.sizes, offsets=table.offsets, indices=table.indices, weights=table.weights,
                            num_stencils=n, weight_streams=lambda nw: [table.weights])
        lib = oref.lib()
        impls = [("cpu", 1)]
        if lib.ref_has_openmp():
            impls.append(("omp", threads))
        for impl, thr in impls:
            if impl == "omp":
                lib.ref_omp_set_threads(thr)
            oref.eval_stencils(src.reshape(-1), (0, L, L), [dst.reshape(-1)], [(0, L, L)], t, impl=impl)   # warm
            ts = []
            for f in range(frames):
                s = frame_primvars(mesh, f + 1)
                t0 = time.perf_counter()
                oref.eval_stencils(s.reshape(-1), (0, L, L), [dst.reshape(-1)], [(0, L, L)], t, impl=impl)
                ts.append(time.perf_counter() - t0)
            res[impl] = {"ms_per_frame": 1e3 * float(np.mean(ts)), "verts_per_s": n / float(np.mean(ts)), "cores": thr,
                         "kind": "reference"}
    else:
        from oracle import oracle
        ts = []
        for f in range(max(1, frames // 2)):
            s = frame_primvars(mesh, f + 1)
            t0 = time.perf_counter()
            oracle.eval_stencils(s.reshape(-1), (0, L, L), [dst.reshape(-1)], [(0, L, L)], table.sizes, table.offsets,
                                 table.indices, [table.weights])
            ts.append(time.perf_counter() - t0)
        res["port"] = {"ms_per_frame": 1e3 * float(np.mean(ts)), "verts_per_s": n / float(np.mean(ts)), "cores": 1, "kind": "port"}
    return res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mesh, table = build_workload()
    threads = os.cpu_count() or 1
    probe = cpu_reference_frames(mesh, table, 1, threads)
    best = max(probe, key=lambda k: probe[k]["verts_per_s"])
    from oracle import ref as oref
    n = table.num_stencils
    dst = np.zeros((n, L), np.float32)
    from types import SimpleNamespace
    t = SimpleNamespace(sizes=table.sizes, offsets=table.offsets, indices=table.indices, weights=table.weights,
                        num_stencils=n, weight_streams=lambda nw: [table.weights])

    # bounded sample: full frames unless K of them would take more than ~2 minutes, then a leading row range
    est = probe[best]["ms_per_frame"] * 1e-3
    rows = n if args.steps * est <= 120.0 else max(100_000, int(n * 120.0 / (args.steps * est)))

    def step(f):
        s = frame_primvars(mesh, f)
        t0 = time.perf_counter()
        if best == "port":
            from oracle import oracle
            oracle.eval_stencils(s.reshape(-1), (0, L, L), [dst.reshape(-1)], [(0, L, L)], table.sizes, table.offsets,
                                 table.indices, [table.weights], 0, rows)
        else:
            oref.eval_stencils(s.reshape(-1), (0, L, L), [dst.reshape(-1)], [(0, L, L)], t, 0, rows, impl=best)
        return time.perf_counter() - t0
    for w in range(args.warmup):
        step(w)
    total = sum(step(args.warmup + k) for k in range(args.steps))
    value = rows * args.steps / total
    cores = probe[best]["cores"]
    # the reference's OpenMP kernel does not scale linearly (false sharing on the result rows): thread sweep
    sweep = {}
    if oref.available() and oref.lib().ref_has_openmp():
        s0 = frame_primvars(mesh, 1)
        thr = 1
        while thr <= threads:
            oref.lib().ref_omp_set_threads(thr)
            oref.eval_stencils(s0.reshape(-1), (0, L, L), [dst.reshape(-1)], [(0, L, L)], t, 0, rows, impl="omp")
            t0 = time.perf_counter()
            oref.eval_stencils(s0.reshape(-1), (0, L, L), [dst.reshape(-1)], [(0, L, L)], t, 0, rows, impl="omp")
            sweep[str(thr)] = round(rows / (time.perf_counter() - t0) / 1e6, 2)
            thr *= 2
        oref.lib().ref_omp_set_threads(threads)
    line = {
        "impl": "reference", "metric": "refined_verts_per_sec_EvalStencils", "value": value, "unit": "verts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": n, "elements": table.num_elements, "control_verts": table.num_control_verts,
                   "primvar_floats": L},
        "cpu_baseline": {"value": value, "unit": "verts/s", "cores": cores, "kind": probe[best]["kind"],
                         "sample": f"{args.steps} steps of rows [0,{rows}) of {n} with Osd::{'OmpEvaluator' if best == 'omp' else 'CpuEvaluator'}"
                                   f" ({best}); probe of all evaluators: "
                                   + ", ".join(f"{k}={v['verts_per_s'] / 1e6:.1f} Mverts/s@{v['cores']}thr" for k, v in probe.items()),
                         "omp_thread_sweep_Mverts_per_s": sweep},
        "e2e": {"value": value, "unit": "verts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------- EvalPatches section --
def bench_eval_patches(mesh, torch, osd, capi, log, n=10_000_000, iters=20, world=1, rank=0, strong=False):
    """BASELINE config 4 shape of work on the synthetic torus: 10 M PatchCoords on 100 000 regular bicubic patches,
    P + 1st + 2nd derivatives of xyz interleaved in one 18-float buffer (glEvalLimit layout), random and patch-sorted
    coordinate order; device-resident, CUDA events.  Algorithmic bytes = n * (20 + 6*12).  Plus the reference's CPU
    evaluators on the first 1 M of the same coordinates.
    N > 1: EvalPatches shards by PatchCoord range with replicated tables and no data-path collective --
    weak: every rank evaluates n coordinates on its own mesh; strong: the n coordinates are cut into N ranges
    (shard.coord_plan).  Time = max over ranks, pts/s = all ranks' coordinates / that time."""
    from opensubdiv_b200 import synth, shard
    D = osd.BufferDescriptor
    n_total = n if (strong or world == 1) else n * world
    lo, hi = (0, n)
    if strong and world > 1:
        lo, hi = shard.coord_plan(n, world, rank).ranges[rank]
    ptab = synth.torus_patch_table(mesh)
    pt = osd.B200PatchTable.Create(ptab)
    src = torch.from_numpy(frame_primvars(mesh, 1)[:, :3].copy()).cuda()
    nl = hi - lo
    out = torch.empty((max(nl, 1), 18), device="cuda")
    args = []
    for k in range(6):
        args += [out, D(3 * k, 3, 18)]
    peak, _ = measured_peak()
    res = {"workload": "torus_400x250_regular_patches_10M_coords_xyz_P+D1+D2", "coords": n_total, "coords_per_gpu": nl,
           "patches": len(mesh.faces), "algorithmic_bytes": n_total * 92,
           "sharding": "none (1 GPU)" if world == 1 else
           ("PatchCoord ranges of one coordinate set, tables replicated" if strong else
            "every rank evaluates its own coordinate set on its own mesh, tables replicated")}
    coords_by_order = {}
    for order, sort in (("random", False), ("sorted_by_patch", True)):
        coords = synth.random_patch_coords(len(mesh.faces), n, seed=2024, sort_by_patch=sort)
        coords_by_order[order] = coords
        pc = torch.from_numpy(np.ascontiguousarray(coords[lo:hi]).view(np.uint8)).cuda()
        for _ in range(3):
            assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, nl, pc, pt, None)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, nl, pc, pt, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            tms = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
            ms = float(tms.item())
        res[order] = {"ms": ms, "pts_per_s": n_total / (ms * 1e-3), "GBps": n_total * 92 / (ms * 1e-3) / 1e9,
                      "frac_of_measured_hbm_peak": n_total * 92 / (ms * 1e-3) / 1e9 / (peak * world)}
        del pc
    if world > 1:
        return res
    try:
        from oracle import ref as oref
        if oref.available():
            m = 1_000_000
            sel = np.ascontiguousarray(coords_by_order["random"][:m])
            tri = oref.PatchTriple(ptab.vertex.arrays, ptab.vertex.indices, ptab.vertex.params)
            srcn = np.ascontiguousarray(frame_primvars(mesh, 1)[:, :3])
            outs = [np.zeros((m, 3), np.float32) for _ in range(6)]
            cpu = {}
            for impl, thr in (("cpu", 1), ("omp", os.cpu_count() or 1)):
                if impl == "omp":
                    if not oref.lib().ref_has_openmp():
                        continue
                    oref.lib().ref_omp_set_threads(thr)
                t0 = time.perf_counter()
                oref.eval_patches(srcn.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in outs], [(0, 3, 3)] * 6, sel, tri, impl=impl)
                dt = time.perf_counter() - t0
                cpu[impl] = {"pts_per_s": m / dt, "cores": thr, "sample": f"{m} of the random coordinates, 1 call"}
            res["cpu_baseline"] = cpu
    except Exception as exc:
        res["cpu_baseline"] = {"error": str(exc)}
    res["find_patches"] = bench_find_patches(torch, osd, n, iters)
    return res


def bench_find_patches(torch, osd, n, iters):
    """(ptexFace, s, t) -> Osd::PatchCoord on the device (B200PatchMap) against Far::PatchMap::FindPatch on
    one host core (its API is one sample per call).  Table: 60 tiled copies of regression shape catmark_car, adaptive
    level 3, Gregory end caps (1.31 M patches, depth 0-3), built by the reference compiled under oracle/_ref.
    Algorithmic bytes = n * (12 in + 20 out); the quadtree (a few MB) stays in L2."""
    try:
        from oracle import ref as oref
        if not oref.available():
            return {"skipped": "oracle/_ref/libosdref.so not present (the adaptive table comes from the reference's factories)"}
        m = oref.Mesh.from_shape_tiled("catmark_car", 60)
        ptab = m.patch_table(3, end_cap="gregory", fvar=False, inf_sharp=True, legacy_sharp_corner=False)
        pm = osd.B200PatchMap.Create(ptab)
        rng = np.random.default_rng(2024)
        face = rng.integers(0, m.num_ptex_faces, n).astype(np.int32)
        s, t = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
        df, ds, dt_ = (torch.from_numpy(x).cuda() for x in (face, s, t))
        pc = torch.empty(n * 5, dtype=torch.int32, device="cuda")
        found = torch.zeros(1, dtype=torch.int32, device="cuda")
        for _ in range(3):
            assert pm.FindPatches(n, df, ds, dt_, pc, found)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            pm.FindPatches(n, df, ds, dt_, pc, found)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        peak, _ = measured_peak()
        k = 2_000_000
        t0 = time.perf_counter()
        want = m.find_patches(ptab, face[:k], s[:k], t[:k])
        cpu_dt = time.perf_counter() - t0
        same = bool(np.array_equal(pc[:k * 5].cpu().numpy(), np.ascontiguousarray(want).view(np.int32)))
        return {"workload": "catmark_car_x60_adaptive_L3_gregory_10M_samples", "samples": n, "patches": len(ptab.vertex.params),
                "max_depth": pm.GetMaxDepth(), "tree_nodes": pm.GetNumNodes(), "found": int(found.item()),
                "ms": ms, "samples_per_s": n / (ms * 1e-3), "algorithmic_bytes": n * 32,
                "GBps": n * 32 / (ms * 1e-3) / 1e9, "frac_of_measured_hbm_peak": n * 32 / (ms * 1e-3) / 1e9 / peak,
                "bit_identical_to_reference_on_sample": same,
                "cpu_baseline": {"samples_per_s": k / cpu_dt, "cores": 1, "kind": "reference",
                                 "sample": f"Far::PatchMap::FindPatch on the first {k} samples"}}
    except Exception as exc:
        return {"error": str(exc)}


def bench_incumbent_cuda(mesh, table, torch, osd):
    """the reference's own CUDA backend kernels (osd/cudaKernel.cu, unmodified, compiled for sm_100a into
    oracle/_ref/libosdcudaref.so) on the same B200 and the same device buffers: config 2 stencils with L = 3 (its tuned
    path) and L = 6 (its generic path), and EvalPatches with 1st + 2nd derivatives.  A reported baseline, like cpu_baseline."""
    from oracle import cuda_ref
    from opensubdiv_b200 import synth
    if not cuda_ref.available():
        return {"skipped": "oracle/_ref/libosdcudaref.so not present"}
    D = osd.BufferDescriptor
    ncv, n = table.num_control_verts, table.num_stencils
    tbl = osd.B200StencilTable.Create(table)
    res = {"kind": "reference CudaEvaluator kernels (osd/cudaKernel.cu), -arch=sm_100a, same box, same buffers"}

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for L in (3, 6):
        pv = frame_primvars(mesh, 1)[:, :L].copy()
        ours = torch.zeros((ncv + n, L), device="cuda")
        ours[:ncv] = torch.from_numpy(pv).cuda()
        theirs = ours.clone()
        args_ref = (theirs.data_ptr(), theirs.data_ptr() + ncv * L * 4, L, L, L, tbl.GetSizesBuffer(), tbl.GetOffsetsBuffer(),
                    tbl.GetIndicesBuffer(), tbl.GetWeightsBuffer(), 0, n)
        ms_ref = timed(lambda: cuda_ref.eval_stencils(*args_ref), 5)
        ms_ours = timed(lambda: osd.B200Evaluator.EvalStencils(ours, D(0, L, L), ours, D(ncv * L, L, L), tbl), 20)
        diff = float((ours[ncv:] - theirs[ncv:]).abs().max().item())
        res[f"eval_stencils_cfg2_L{L}"] = {"reference_cuda_ms": ms_ref, "reference_cuda_verts_per_s": n / (ms_ref * 1e-3),
                                           "b200osd_ms": ms_ours, "b200osd_verts_per_s": n / (ms_ours * 1e-3),
                                           "max_abs_diff": diff}
        del ours, theirs
    # EvalPatches: 2 M random coords on the torus patches, P + D1 + D2 interleaved (18 floats)
    m = 2_000_000
    ptab = synth.torus_patch_table(mesh)
    pt = osd.B200PatchTable.Create(ptab)
    coords = synth.random_patch_coords(len(mesh.faces), m, seed=2024)
    pc = torch.from_numpy(coords.view(np.uint8)).cuda()
    src = torch.from_numpy(frame_primvars(mesh, 1)[:, :3].copy()).cuda()
    ours = torch.zeros((m, 18), device="cuda")
    theirs = torch.zeros((m, 18), device="cuda")
    a = []
    for k in range(6):
        a += [ours, D(3 * k, 3, 18)]
    ms_ref = timed(lambda: cuda_ref.eval_patches(src.data_ptr(), [theirs.data_ptr() + 12 * k for k in range(6)], 3, 3, [18] * 6,
                                                 m, pc.data_ptr(), pt.GetPatchArrayBuffer(), pt.GetPatchIndexBuffer(),
                                                 pt.GetPatchParamBuffer()), 3)
    ms_ours = timed(lambda: osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *a, m, pc, pt, None), 20)
    res["eval_patches_2M_coords_P+D1+D2"] = {"reference_cuda_ms": ms_ref, "reference_cuda_pts_per_s": m / (ms_ref * 1e-3),
                                            "b200osd_ms": ms_ours, "b200osd_pts_per_s": m / (ms_ours * 1e-3),
                                            "max_abs_diff": float((ours - theirs).abs().max().item())}
    return res


# --------------------------------------------------------------------------------- B200 arm --
def run_b200_arm(args):
    import faulthandler
    # a run must never hang the box: dump every thread's stack and leave after 5 minutes (a first `import torch` on a
    # fresh box alone can take a minute)
    faulthandler.dump_traceback_later(300, exit=True)
    import torch
    import torch.distributed as dist
    import opensubdiv_b200 as osd
    from opensubdiv_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL kernels on a high-priority stream: their few CTAs are dispatched as soon as an SM slot frees up instead
        # of queueing behind the 25 000-block evaluation grid (which would serialise broadcast and kernel)
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
        except Exception:
            opts = None
        if opts is not None:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    D = osd.BufferDescriptor
    lib = capi.lib()

    from opensubdiv_b200 import shard
    mesh, table = build_workload()
    ncv = table.num_control_verts
    strong = (args.scaling == "strong") and world > 1
    if strong:
        # one mesh, rows cut into `world` contiguous ranges balanced on stencil elements; every rank needs all control points
        plan = shard.ShardPlan.for_table(table.sizes, world, rank, align=2048)
        alg_bytes_total = table.algorithmic_bytes(1, L, L)
        table = shard.local_table(table, plan)
        meshes_in_scene, my_mesh = 1, 0
        total_rows = sum(b - a for a, b in plan.ranges)
    else:
        # `world` meshes; rank r owns all rows of mesh r; the scene's control points are replicated every frame
        alg_bytes_total = table.algorithmic_bytes(1, L, L) * world
        meshes_in_scene, my_mesh = world, rank
        total_rows = table.num_stencils * world
    n = table.num_stencils
    t0 = time.time()
    tbl = osd.B200StencilTable.Create(table)
    assert tbl is not None, capi.last_error()
    log(f"[bench] rank {rank}: B200StencilTable of {n} rows built in {time.time() - t0:.1f}s")

    # vertex buffer = [ control block A | control block B | this rank's refined rows ]; A/B are the two halves of the
    # double-buffered per-frame broadcast (frame f lives in block f % 2)
    scene_cv = ncv * meshes_in_scene
    # ... and two refined regions so that (host-buffer path) the D2H read-back of frame f overlaps frame f+1's kernel
    vb = osd.B200VertexBuffer.Create(L, 2 * scene_cv + 2 * n)
    assert vb is not None, capi.last_error()
    vt = vb.as_tensor()
    blocks = [vt[:scene_cv], vt[scene_cv:2 * scene_cv]]
    src_descs = [D((b * scene_cv + my_mesh * ncv) * L, L, L) for b in (0, 1)]
    dst_vertex = [2 * scene_cv, 2 * scene_cv + n]
    dst_descs = [D(v * L, L, L) for v in dst_vertex]
    bc = shard.FrameBroadcaster(blocks, root=0)

    frames = [torch.from_numpy(np.tile(frame_primvars(mesh, f), (meshes_in_scene, 1))).pin_memory() for f in range(4)]
    host_out = [torch.empty((n, L), dtype=torch.float32).pin_memory() for _ in range(2)]
    stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    kernel_done = [torch.cuda.Event(), torch.cuda.Event()]
    d2h_done = [torch.cuda.Event(), torch.cuda.Event()]
    for e in d2h_done:
        e.record(stream)
    for b in (0, 1):                                       # device-resident control points for the `value` measurement
        vb.UpdateData(frames[b], b * scene_cv, scene_cv)
    torch.cuda.synchronize()

    # Frames are pipelined: the broadcast of frame f+1 is posted BEFORE frame f's kernel is enqueued, so on the side
    # stream it only waits for the kernel that last read its buffer (frame f-1) and travels while frame f is evaluated.
    state = {"f": 0, "posted": -1}

    def advance(e2e):
        f = state["f"]
        for g in (f, f + 1):                                   # prologue posts f, steady state posts only f+1
            if g > state["posted"]:
                if e2e and rank == 0:                          # host-buffer path: this frame's control points H2D (root)
                    vb.UpdateData(frames[g % len(frames)], (g % 2) * scene_cv, scene_cv)
                bc.post(g)
                state["posted"] = g
        bc.wait(f)
        r = f % 2 if e2e else 0
        if e2e:
            stream.wait_event(d2h_done[r])                     # refined region r: read-back of frame f-2 has finished
        ok = osd.B200Evaluator.EvalStencils(vb, src_descs[f % 2], vb, dst_descs[r], tbl)
        assert ok
        bc.release(f)
        if e2e:                                                # D2H of this rank's refined vertices on the copy stream
            kernel_done[r].record(stream)
            copy_stream.wait_event(kernel_done[r])
            vb.ReadData(host_out[r], dst_vertex[r], n, deviceContext=copy_stream)
            d2h_done[r].record(copy_stream)
        state["f"] = f + 1

    def step_device(_):
        """Device-resident step: control points already in HBM on the root; (N>1: per-frame broadcast, overlapped
        with the previous frame's kernel) + EvalStencils of this rank's rows."""
        advance(False)

    def step_e2e(_):
        """Host-buffer step through the C ABI: H2D control points (root), replicate, evaluate, D2H refined vertices."""
        advance(True)

    # N > 1: the per-frame Python cost of torch.distributed.broadcast (tens of microseconds) is of the order of the
    # kernel itself, so two consecutive frames (one per control block) are captured ONCE into a CUDA graph --
    # kernel(block b) runs concurrently with broadcast(block 1-b) -- and the timed loop replays it.
    graph = None
    if world > 1 and args.graph:
        try:
            for b in (0, 1):                                   # both blocks valid everywhere before the first replay
                dist.broadcast(blocks[b], src=0)
            torch.cuda.synchronize()
            comm = torch.cuda.Stream(priority=-1)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cur = torch.cuda.current_stream()
                for b in (0, 1):
                    comm.wait_stream(cur)                      # block 1-b's last reader (previous kernel) has finished
                    with torch.cuda.stream(comm):
                        dist.broadcast(blocks[1 - b], src=0)
                    ok = osd.B200Evaluator.EvalStencils(vb, src_descs[b], vb, dst_descs[0], tbl)
                    assert ok
                    cur.wait_stream(comm)                      # next kernel reads block 1-b
            graph = g
            log(f"[bench] rank {rank}: frame pair captured into a CUDA graph")
        except Exception as exc:
            graph = None
            log(f"[bench] rank {rank}: CUDA graph capture failed ({exc}); eager pipeline")
            torch.cuda.synchronize()

    pending = {"half": 0}

    def step_graph(_):
        """One frame = half a replay of the captured pair (replayed on every second call)."""
        if pending["half"] == 0:
            graph.replay()
        pending["half"] ^= 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_device(0)
    torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        for w in range(warmup):
            step_fn(w)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.b200osd_reset_launch_count()
        e0.record(stream)
        for k in range(steps):
            step_fn(warmup + k)
        for e in d2h_done:                                     # read-backs still in flight belong to the timed region
            stream.wait_event(e)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.b200osd_launch_count()
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    log(f"[bench] rank {rank}: timing {args.steps} device-resident steps")
    if graph is not None and args.steps % 2 == 0:
        pending["half"] = 0
        ms_dev, _ = timed(step_graph, args.steps, args.warmup + (args.warmup % 2))
        launches = args.steps                                  # one sell_kernel per frame inside the replayed graph
    else:
        ms_dev, launches = timed(step_device, args.steps, args.warmup)
    log(f"[bench] rank {rank}: device-resident done ({ms_dev / args.steps:.4f} ms/step); timing host-buffer steps")
    ms_e2e, _ = timed(step_e2e, max(3, min(args.steps, 20)), 3)
    log(f"[bench] rank {rank}: host-buffer steps done")
    e2e_steps = max(3, min(args.steps, 20))
    # the timed regions are milliseconds long: keep the same step running ~1 s more (untimed) so that the 50 ms
    # nvidia-smi sampler sees the clocks this kernel actually runs at under load
    # The number of extra steps must be IDENTICAL on every rank (each step posts a broadcast; a wall-clock loop lets ranks
    # disagree by one batch and leaves unmatched collectives behind): derive it from the max-reduced step time.
    batches = int(min(400, max(1, round(1.0 / max(50 * (ms_dev / args.steps) * 1e-3, 1e-4)))))
    for _ in range(batches):
        for k in range(50):
            step_device(k)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = ms_dev / args.steps
    value = total_rows * args.steps / (ms_dev * 1e-3)
    e2e_value = total_rows * e2e_steps / (ms_e2e * 1e-3)
    alg_bytes = alg_bytes_total // world                      # per GPU, per launch
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9

    # Second half of BASELINE.json's metric (limit pts/s, EvalPatches): reported in "eval_patches" (sharded by PatchCoord range at N > 1).
    patches = None
    if not args.no_patches:
        if world == 1:
            try:
                patches = bench_eval_patches(mesh, torch, osd, capi, log)
            except Exception as exc:
                patches = {"error": str(exc)}
        elif args.shard_patches:
            # collectives inside: every rank must take the same path, so no exception is swallowed here
            patches = bench_eval_patches(mesh, torch, osd, capi, log, world=world, rank=rank, strong=strong)
        else:
            patches = {"skipped": "N > 1: run with --shard-patches for EvalPatches sharded by PatchCoord range "
                                  "(measured at N=2: profiles/r01_bench_n2.json)"}

    incumbent = None
    if world == 1 and not args.no_patches:
        try:
            incumbent = bench_incumbent_cuda(mesh, table, torch, osd)
        except Exception as exc:
            incumbent = {"error": str(exc)}

    if rank == 0:
        cpu = None
        if world == 1:                                          # the CPU baseline is a rank-0, N=1 report
            try:
                probe = cpu_reference_frames(mesh, table, 4, os.cpu_count() or 1)
                best = max(probe, key=lambda k: probe[k]["verts_per_s"])
                cpu = {"value": probe[best]["verts_per_s"], "unit": "verts/s", "cores": probe[best]["cores"],
                       "kind": probe[best]["kind"],
                       "sample": "4 full frames of the same 6.4 M-row table per evaluator; "
                                 + ", ".join(f"{k}: {v['ms_per_frame']:.0f} ms/frame @{v['cores']} thr" for k, v in probe.items())}
            except Exception as exc:      # the baseline is a report, never a reason to lose the GPU number
                cpu = {"value": None, "unit": "verts/s", "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
        else:
            cpu = {"value": None, "unit": "verts/s", "cores": 0, "kind": "reference",
                   "sample": "reported at N=1 only (run bench.py --gpus 1 or --impl reference)"}
        line = {
            "metric": "refined_verts_per_sec_EvalStencils", "value": value, "unit": "verts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": n, "elements_per_gpu": table.num_elements,
                       "control_verts_per_mesh": ncv, "primvar_floats": L, "parallelism": f"row-range x{world}",
                       "launch": "eager" if graph is None else "CUDA graph of 2 frames (kernel || broadcast of the other control block)",
                       "exchange": "none (1 GPU)" if world == 1 else
                       f"per-frame NCCL broadcast of {scene_cv * L * 4} B of control points from rank 0, double-buffered on a side stream",
                       "l2_policy": "inputs larger than L2 (table streams 0.7 GB/step vs 126 MB L2); no flush needed",
                       "stencil_variant": lib.b200osd_get_stencil_variant(),
                       "bucketed_stream_bytes": tbl.GetStreamBytes(1)},
            "e2e": {"value": e2e_value, "unit": "verts/s", "h2d_bytes_per_step": int(scene_cv * L * 4),
                    "d2h_bytes_per_step": int(n * L * 4), "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "frac_of_nominal_8TBps": achieved / 8000.0,
                         "note": "per GPU; device time per step = one sell_kernel launch (+ the overlapped broadcast when N > 1)"},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "eval_patches": patches,
            "incumbent_cuda": incumbent,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing NCCL down: destroying a process group that a captured graph still references can hang
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="stencil kernel variant (0 = auto)")
    ap.add_argument("--no-patches", action="store_true", help="skip the EvalPatches section of the report")
    ap.add_argument("--shard-patches", action="store_true",
                    help="N > 1: also time EvalPatches sharded by PatchCoord range across the ranks")
    ap.add_argument("--graph", action="store_true",
                    help="N > 1: replay a CUDA graph of two frames instead of the eager pipeline (verified at N=2 only)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = N meshes (default), strong = one mesh cut into N row ranges")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.variant:
        from opensubdiv_b200 import capi
        capi.lib().b200osd_set_stencil_variant(args.variant)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
