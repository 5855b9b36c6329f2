"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run HERE (needs /root/reference compiled in place: `make -C oracle/ref`):   python tests/golden/make_golden.py

Every fixture holds reference-built tables (Far::StencilTable / LimitStencilTable / PatchTable flattened by
Osd::CpuPatchTable, PatchCoords from Far::PatchMap::FindPatch), the inputs, and the outputs of the reference's own
Osd::CpuEvaluator on them.  catmark_cube_L4 additionally carries the reference's hbr_regression baseline positions
(regression/hbr_regression/baseline/catmark_cube_level3.obj = 4 levels of subdivision).
The fixtures travel to the GPU box, where the reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref  # noqa: E402

REF_ROOT = "/root/reference"


def table_dict(prefix, st):
    d = {prefix + "ncv": np.int32(st.num_control_verts), prefix + "sizes": st.sizes, prefix + "offsets": st.offsets,
         prefix + "indices": st.indices, prefix + "weights": st.weights}
    for k in ("du", "dv", "duu", "duv", "dvv"):
        a = getattr(st, k)
        if a is not None:
            d[prefix + k] = a
    return d


def primvar(mesh, extra=0, seed=7):
    """xyz positions (+ `extra` pseudo-random floats per vertex)."""
    pos = mesh.positions
    if extra == 0:
        return pos.copy()
    rng = np.random.default_rng(seed)
    return np.concatenate([pos, rng.standard_normal((len(pos), extra)).astype(np.float32)], axis=1)


def run_stencils(src, st, L, nw=1):
    """Same-buffer layout of Osd::Mesh::Refine (osd/mesh.h:505-519): [control | refined]."""
    ncv, n = st.num_control_verts, st.num_stencils
    outs = []
    buf = np.zeros((ncv + n, L), np.float32)
    buf[:ncv] = src
    flat = buf.reshape(-1)
    if nw == 1:
        assert ref.eval_stencils(flat, (0, L, L), [flat], [(ncv * L, L, L)], st)
        return [buf[ncv:].copy()]
    for _ in range(nw):
        outs.append(np.zeros((n, L), np.float32))
    assert ref.eval_stencils(flat, (0, L, L), [o.reshape(-1) for o in outs], [(0, L, L)] * nw, st)
    return outs


def read_obj_positions(path):
    pts = []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                pts.append([float(x) for x in line.split()[1:4]])
    return np.asarray(pts, dtype=np.float32)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-40s %8.1f KB" % (name, os.path.getsize(path) / 1024.0))


def make_cube():
    m = ref.Mesh.from_shape("catmark_cube").refine_uniform(4)
    last = m.stencil_table(intermediate_levels=False)
    alll = m.stencil_table(intermediate_levels=True)
    src = primvar(m)
    d = {"src": src, "hbr_level3": read_obj_positions(
        os.path.join(REF_ROOT, "regression/hbr_regression/baseline/catmark_cube_level3.obj"))}
    d.update(table_dict("last_", last))
    d.update(table_dict("all_", alll))
    d["last_out"] = run_stencils(src, last, 3)[0]
    d["all_out"] = run_stencils(src, alll, 3)[0]
    save("catmark_cube_L4", **d)


def make_stencil_shapes():
    for shape, level, extra in (("loop_icosahedron", 3, 0), ("catmark_pyramid_creases0", 3, 1), ("catmark_pole64", 2, 3),
                                ("bilinear_cube", 3, 0), ("catmark_cube_corner2", 3, 5), ("loop_cube_creases1", 2, 0),
                                ("catmark_tent_creases0", 3, 0), ("catmark_car", 1, 0)):
        m = ref.Mesh.from_shape(shape).refine_uniform(level)
        st = m.stencil_table(intermediate_levels=True)
        src = primvar(m, extra)
        d = {"src": src}
        d.update(table_dict("t_", st))
        d["out"] = run_stencils(src, st, src.shape[1])[0]
        # varying stencils of the same refinement (bilinear, sizes 1/2/4)
        vst = m.stencil_table(mode="varying", intermediate_levels=True)
        d.update(table_dict("v_", vst))
        d["v_out"] = run_stencils(src, vst, src.shape[1])[0]
        save("stencils_" + shape, **d)


def sample_locations(mesh, per_face, seed):
    rng = np.random.default_rng(seed)
    nf = mesh.num_ptex_faces
    face = np.repeat(np.arange(nf, dtype=np.int32), per_face)
    s = rng.random(len(face), dtype=np.float32)
    t = rng.random(len(face), dtype=np.float32)
    if mesh.reg_face_size == 3:                         # keep inside the triangle (glStencilViewer.cpp:364-371)
        flip = (s + t) >= 1.0
        s = np.where(flip, 1.0 - s, s).astype(np.float32)
        t = np.where(flip, 1.0 - t, t).astype(np.float32)
    # a few exact corners / edges / centres
    s[::17] = 0.0
    t[::19] = 0.0
    s[5::23] = 0.5
    if mesh.reg_face_size == 3:
        t = np.minimum(t, 1.0 - s).astype(np.float32)
    return face, s, t


def make_limit():
    for shape, level in (("catmark_cube_creases0", 3), ("loop_cube", 3), ("catmark_gregory_test2", 3), ("catmark_torus", 2)):
        m = ref.Mesh.from_shape(shape).refine_adaptive(level)
        face, s, t = sample_locations(m, 12, 11)
        st = m.limit_stencil_table(face, s, t, first=True, second=True)
        src = primvar(m)
        d = {"src": src, "face": face, "s": s, "t": t}
        d.update(table_dict("t_", st))
        outs = run_stencils(src, st, 3, nw=6)
        for k, o in zip(("p", "du", "dv", "duu", "duv", "dvv"), outs):
            d["out_" + k] = o
        save("limit_" + shape, **d)


def triple_dict(prefix, tr):
    return {prefix + "arrays": tr.arrays, prefix + "indices": tr.indices, prefix + "params": tr.params}


def make_patches():
    cases = (
        # shape, level, endcap, fvar
        ("catmark_cube_creases0", 3, "gregory", False),
        ("catmark_gregory_test2", 3, "gregory", False),
        ("catmark_gregory_test4", 2, "bspline", False),
        ("catmark_fvar_bound1", 3, "gregory", True),
        ("catmark_edgecorner", 3, "gregory", False),
        ("catmark_pyramid", 2, "bilinear", False),
        ("loop_icosahedron", 3, "gregory", False),
        ("loop_cube_creases0", 2, "gregory", False),
        ("loop_triangle_edgeonly", 3, "gregory", False),
        ("catmark_car", 2, "gregory", True),
    )
    for shape, level, endcap, fvar in cases:
        m = ref.Mesh.from_shape(shape)
        pt = m.patch_table(level, end_cap=endcap, fvar=fvar, fvar_legacy_linear=False, inf_sharp=True,
                           legacy_sharp_corner=False, refine_first=True)
        st = m.stencil_table(intermediate_levels=True, patch_table=pt)
        src0 = primvar(m)
        ncv, n = st.num_control_verts, st.num_stencils
        vb = np.zeros((ncv + n, 3), np.float32)
        vb[:ncv] = src0
        assert ref.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st)
        face, s, t = sample_locations(m, 24 if ncv < 200 else 2, 5)
        pc = m.find_patches(pt, face, s, t)
        keep = pc["arrayIndex"] >= 0
        pc = pc[keep]
        d = {"src0": src0, "vb": vb, "coords": pc}
        d.update(table_dict("st_", st))
        d.update(triple_dict("vtx_", pt.vertex))
        outs = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
        assert ref.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in outs], [(0, 3, 3)] * 6, pc, pt.vertex)
        for k, o in zip(("p", "du", "dv", "duu", "duv", "dvv"), outs):
            d["out_" + k] = o
        # varying: same refined buffer evaluated through the varying (linear) patches
        if pt.varying is not None:
            vst = m.stencil_table(mode="varying", intermediate_levels=True, patch_table=pt)
            vvb = np.zeros((vst.num_control_verts + vst.num_stencils, 3), np.float32)
            vvb[:ncv] = src0
            assert ref.eval_stencils(vvb.reshape(-1), (0, 3, 3), [vvb.reshape(-1)], [(ncv * 3, 3, 3)], vst)
            d.update(triple_dict("var_", ref.PatchTriple(pt.varying.arrays, pt.varying.indices, pt.vertex.params)))
            d["var_vb"] = vvb
            vo = [np.zeros((len(pc), 3), np.float32) for _ in range(3)]
            assert ref.eval_patches(vvb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in vo], [(0, 3, 3)] * 3, pc,
                                    ref.PatchTriple(pt.varying.arrays, pt.varying.indices, pt.vertex.params))
            for k, o in zip(("p", "du", "dv"), vo):
                d["var_out_" + k] = o
        if fvar and pt.fvar:
            fst = m.stencil_table(mode="fvar", intermediate_levels=True, patch_table=pt)
            uv0 = m.uvs
            nfv = fst.num_control_verts
            fvb = np.zeros((nfv + fst.num_stencils, 2), np.float32)
            fvb[:nfv] = uv0[:nfv]
            assert ref.eval_stencils(fvb.reshape(-1), (0, 2, 2), [fvb.reshape(-1)], [(nfv * 2, 2, 2)], fst)
            d.update(table_dict("fst_", fst))
            d.update(triple_dict("fvar_", pt.fvar[0]))
            d["fvar_vb"] = fvb
            fo = [np.zeros((len(pc), 2), np.float32) for _ in range(6)]
            assert ref.eval_patches(fvb.reshape(-1), (0, 2, 2), [o.reshape(-1) for o in fo], [(0, 2, 2)] * 6, pc, pt.fvar[0])
            for k, o in zip(("p", "du", "dv", "duu", "duv", "dvv"), fo):
                d["fvar_out_" + k] = o
        save("patches_" + shape, **d)


PATCH_MAP_CASES = (
    # shape, isolation level, end cap
    ("catmark_cube", 4, "gregory"),
    ("catmark_car", 2, "gregory"),
    ("catmark_nonquads", 3, "gregory"),          # non-quad base faces: ptex sub-faces, root depth 1
    ("catmark_hole_test2", 3, "gregory"),        # hole faces: FindPatch returns NULL
    ("catmark_pole64", 5, "bspline"),
    ("catmark_edgecorner", 6, "gregory"),
    ("loop_icosahedron", 4, "gregory"),          # triangular domain with rotated sub-triangles
    ("loop_cube_creases0", 3, "gregory"),
    ("loop_triangle_edgecorner", 5, "gregory"),
)


def patch_map_samples(mesh, n, seed):
    """(face, s, t) locations stressing the quadtree: random interior points, every dyadic boundary down to 2^-7,
    exact 0 / 1, points on the diagonals of triangular domains, and faces outside the table."""
    rng = np.random.default_rng(seed)
    nf = mesh.num_ptex_faces
    face = rng.integers(-2, nf + 2, n).astype(np.int32)
    s = rng.random(n, dtype=np.float32)
    t = rng.random(n, dtype=np.float32)
    k = rng.integers(0, 129, n)
    snap_s, snap_t = rng.random(n) < 0.3, rng.random(n) < 0.3
    s = np.where(snap_s, (k / 128.0), s).astype(np.float32)
    t = np.where(snap_t, (rng.integers(0, 129, n) / 128.0), t).astype(np.float32)
    if mesh.reg_face_size == 3:
        flip = (s + t) > 1.0
        s, t = np.where(flip, 1.0 - s, s).astype(np.float32), np.where(flip, 1.0 - t, t).astype(np.float32)
        diag = rng.random(n) < 0.2                       # exactly on a sub-triangle diagonal s + t = j / 2^m
        m = 2.0 ** rng.integers(0, 6, n)
        j = np.floor(rng.random(n) * m) + 1
        td = (j / m - s).astype(np.float32)
        ok = diag & (td >= 0) & (td <= 1)
        t = np.where(ok, td, t).astype(np.float32)
    return face, s, t


def make_patch_maps():
    for shape, level, endcap in PATCH_MAP_CASES:
        m = ref.Mesh.from_shape(shape)
        pt = m.patch_table(level, end_cap=endcap, fvar=False, inf_sharp=True, legacy_sharp_corner=False, refine_first=True)
        face, s, t = patch_map_samples(m, 6000, 3)
        pc = m.find_patches(pt, face, s, t)
        save("patchmap_" + shape, arrays=pt.vertex.arrays, params=pt.vertex.params,
             triangular=np.int32(m.reg_face_size == 3), num_ptex_faces=np.int32(m.num_ptex_faces),
             face=face, s=s, t=t, coords=pc)


if __name__ == "__main__":
    groups = {"cube": make_cube, "stencils": make_stencil_shapes, "limit": make_limit, "patches": make_patches,
              "patchmap": make_patch_maps}
    for g in (sys.argv[1:] or list(groups)):
        groups[g]()
