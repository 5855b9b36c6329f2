"""The CUDA patch kernel's per-coordinate arithmetic, compiled for the HOST (tests/emu), against the reference's golden
outputs -- lets the kernel maths be checked in this GPU-less container.  Test-only harness; not a product path."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from oracle import oracle
from tests.util import golden, golden_names, triple_from, assert_close

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libpatch_emu.so")
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")


def _lib():
    if os.path.exists("/usr/local/cuda/bin/nvcc") and shutil.which("g++"):
        from tests.emu import build
        build.build()
    if not os.path.exists(EMU):
        pytest.skip("tests/emu/libpatch_emu.so not built and nvcc unavailable")
    L = C.CDLL(EMU)
    L.emu_eval_patches.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4
    return L


def emu_patches(L, src, desc, n_comp, coords, tr, nw):
    outs = [np.zeros((len(coords), n_comp), np.float32) for _ in range(nw)]
    sd = (C.c_int * 3)(*desc)
    dd = (C.c_int * (3 * nw))(*([0, n_comp, n_comp] * nw))
    dptr = (C.c_void_p * nw)(*[o.ctypes.data for o in outs])
    a, ix, pr = (np.ascontiguousarray(x) for x in (tr.arrays, tr.indices, tr.params))
    src = np.ascontiguousarray(src)
    coords = np.ascontiguousarray(coords)
    L.emu_eval_patches(src.ctypes.data, sd, nw, dptr, dd, len(coords), coords.ctypes.data, a.ctypes.data, ix.ctypes.data,
                       pr.ctypes.data)
    return outs


@pytest.mark.parametrize("name", golden_names("patches_"))
@pytest.mark.parametrize("nw", [1, 3, 6])
def test_patch_kernel_math_vs_cpu_evaluator(name, nw):
    L = _lib()
    d = golden(name)
    tr = triple_from(d, "vtx_")
    coords, vb = d["coords"], d["vb"]
    got = emu_patches(L, vb.reshape(-1), (0, 3, 3), 3, coords, tr, nw)
    scl = [np.zeros((len(coords), 3), np.float32) for _ in range(nw)]
    with oracle.abs_mode(2):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in scl], [(0, 3, 3)] * nw, coords, tr.arrays,
                            tr.indices, tr.params)
    for k in range(nw):
        assert_close(got[k], d["out_" + OUT6[k]], scl[k], f"{name} {OUT6[k]}")
    if "fvar_out_p" in d.files and nw == 6:
        ftr = triple_from(d, "fvar_")
        fg = emu_patches(L, d["fvar_vb"].reshape(-1), (0, 2, 2), 2, coords, ftr, 6)
        fs = [np.zeros((len(coords), 2), np.float32) for _ in range(6)]
        with oracle.abs_mode(2):
            oracle.eval_patches(d["fvar_vb"].reshape(-1), (0, 2, 2), [o.reshape(-1) for o in fs], [(0, 2, 2)] * 6, coords,
                                ftr.arrays, ftr.indices, ftr.params)
        for k in range(6):
            assert_close(fg[k], d["fvar_out_" + OUT6[k]], fs[k], f"{name} fvar {OUT6[k]}")


def test_interior_regular_patches_meet_the_plain_bound():
    """Where no weight cancellation exists (interior bicubic B-spline patches) the kernel meets 1e-6 against the plain
    scale sum|w||x| -- the separable evaluation order and FMA are the only differences from the reference."""
    L = _lib()
    d = golden("patches_catmark_car")
    tr = triple_from(d, "vtx_")
    coords, vb = d["coords"], d["vb"]
    f1 = tr.params["field1"][coords["patchIndex"]]
    interior = (((f1 >> 7) & 31) == 0) & (((f1 >> 5) & 1) == 1)
    sel = np.ascontiguousarray(coords[interior])
    assert len(sel) > 1000
    got = emu_patches(L, vb.reshape(-1), (0, 3, 3), 3, sel, tr, 6)
    scl = [np.zeros((len(sel), 3), np.float32) for _ in range(6)]
    with oracle.abs_mode(1):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in scl], [(0, 3, 3)] * 6, sel, tr.arrays,
                            tr.indices, tr.params)
    for k in range(6):
        assert_close(got[k], d["out_" + OUT6[k]][interior], scl[k], f"interior {OUT6[k]}")


def _all_shapes():
    from oracle import ref
    return [s for s in ref.shape_names() if not s.startswith("bilinear")] if ref.available() else []


@pytest.mark.parametrize("shape", _all_shapes())
def test_patch_kernel_math_on_every_regression_shape(shape):
    """The CUDA kernel's arithmetic (host emulation) against the unmodified reference's Osd::CpuEvaluator::EvalPatches on
    EVERY shape of the reference's regression list (regression/osd_regression/main.cpp:250-330): adaptive level 3, Gregory
    end caps, 1 500 random samples, P + D1 + D2, 1e-6 in the conditioned scale; and the emulated device patch map against
    Far::PatchMap::FindPatch, bit-exact.  Skipped where the reference build is absent (the GPU box runs the golden subset)."""
    from oracle import ref
    L = _lib()
    m = ref.Mesh.from_shape(shape)
    # the valence-360 pole costs the reference 100 s of end-cap stencil building at level 3: one level is enough there
    pt = m.patch_table(1 if shape.endswith("pole360") else 3, end_cap="gregory", inf_sharp=True, legacy_sharp_corner=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    ncv, n = st.num_control_verts, st.num_stencils
    vb = np.zeros((ncv + n, 3), np.float32)
    vb[:ncv] = m.positions
    assert ref.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st)
    rng = np.random.default_rng(11)
    k = 1500
    face = rng.integers(0, m.num_ptex_faces, k).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    s[::37], t[::41] = 0.0, 1.0
    if m.reg_face_size == 3:
        flip = s + t >= 1
        s, t = np.where(flip, 1 - s, s).astype(np.float32), np.where(flip, 1 - t, t).astype(np.float32)
    want = m.find_patches(pt, face, s, t)
    pc = np.ascontiguousarray(want[want["arrayIndex"] >= 0])
    assert len(pc) > 0
    x = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    z = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    assert ref.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in x], [(0, 3, 3)] * 6, pc, pt.vertex)
    with oracle.abs_mode(2):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in z], [(0, 3, 3)] * 6, pc, pt.vertex.arrays,
                            pt.vertex.indices, pt.vertex.params)
    got = emu_patches(L, vb.reshape(-1), (0, 3, 3), 3, pc, pt.vertex, 6)
    for kk in range(6):
        assert_close(got[kk], x[kk], z[kk], f"{shape} {OUT6[kk]}")


def synthetic_patch_sweep(ptype):
    """One synthetic patch per (type, boundary mask, depth, sub-patch, rotation) with a sample each, incl. corners and
    edges: (triple, coords, vertex buffer).  Shared with tests/test_gpu_patches.py."""
    from oracle.ref import PATCH_ARRAY_DTYPE, PATCH_PARAM_DTYPE, PATCH_COORD_DTYPE
    from types import SimpleNamespace
    npts = {3: 4, 4: 3, 5: 12, 6: 16, 9: 20, 10: 18}[ptype]
    tri = ptype in (4, 5, 10)
    rng = np.random.default_rng(100 + ptype)
    masks = range(32) if ptype == 5 else (range(16) if ptype == 6 else [0])
    params, coords = [], []
    for mask in masks:
        for trial in range(40):
            depth = int(rng.integers(0, 7))
            nonquad = int(rng.integers(0, 2)) if depth > 0 else 0
            n = 1 << (depth - nonquad)
            u, v = int(rng.integers(0, n)), int(rng.integers(0, n))
            a, b = rng.random(2)
            if tri and a + b > 1:
                a, b = 1 - a, 1 - b
            if trial % 7 == 0:
                a = 0.0
            if trial % 11 == 0:
                b = 0.0
            if trial % 13 == 0:
                a, b = (1.0, 0.0) if tri else (1.0, 1.0)
            if tri and (u + v) >= (1 << depth):
                s, t = ((1 << depth) - u - a) / n, ((1 << depth) - v - b) / n
            else:
                s, t = (u + a) / n, (v + b) / n
            f1 = depth | (nonquad << 4) | (1 << 5) | (mask << 7) | (v << 12) | (u << 22)
            coords.append((0, len(params), 0, np.float32(s), np.float32(t)))
            params.append((0, f1, 0.0))
    P = len(params)
    tr = SimpleNamespace(arrays=np.array([(ptype, ptype, P, 0, npts, 0)], PATCH_ARRAY_DTYPE),
                         indices=(np.arange(P * npts, dtype=np.int32) % (npts * 7)).astype(np.int32),
                         params=np.array(params, PATCH_PARAM_DTYPE))
    pc = np.array(coords, PATCH_COORD_DTYPE)
    vb = rng.standard_normal((npts * 7, 3)).astype(np.float32)
    return tr, pc, vb


@pytest.mark.parametrize("ptype", [3, 4, 5, 6, 9, 10])
def test_patch_kernel_math_every_boundary_mask_depth_rotation(ptype):
    """The kernel's arithmetic against the oracle (itself bit-pinned to OsdEvaluatePatchBasis,
    tests/test_oracle_vs_reference.py) for PatchParam combinations that real tables rarely contain -- all 32 Loop masks, all
    16 B-spline masks, depths 0..6, rotated triangles, non-quad roots, samples on patch corners and edges."""
    L = _lib()
    tr, pc, vb = synthetic_patch_sweep(ptype)
    P = len(pc)
    exp = [np.zeros((P, 3), np.float32) for _ in range(6)]
    scl = [np.zeros((P, 3), np.float32) for _ in range(6)]
    assert oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in exp], [(0, 3, 3)] * 6, pc, tr.arrays, tr.indices, tr.params)
    with oracle.abs_mode(2):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in scl], [(0, 3, 3)] * 6, pc, tr.arrays, tr.indices, tr.params)
    got = emu_patches(L, vb.reshape(-1), (0, 3, 3), 3, pc, tr, 6)
    for k in range(6):
        assert_close(got[k], exp[k], scl[k], f"type {ptype} {OUT6[k]}")
