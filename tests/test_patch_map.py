"""Patch map: the library's host tree builder + the kernel's descent, compiled for the host (tests/emu),
against Far::PatchMap::FindPatch outputs recorded from the reference (tests/golden/patchmap_*.npz) and, where the
reference is compiled here, against fresh calls on more shapes.  Integer work: bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.test_kernel_math_emu import _lib
from tests.util import golden, golden_names

COORD = np.dtype([("arrayIndex", "<i4"), ("patchIndex", "<i4"), ("vertIndex", "<i4"), ("s", "<f4"), ("t", "<f4")])


def emu_find(L, arrays, params, triangular, face, s, t):
    L.emu_patch_map_find.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
    arrays, params = np.ascontiguousarray(arrays), np.ascontiguousarray(params)
    face, s, t = (np.ascontiguousarray(x) for x in (face, s, t))
    out = np.zeros(len(face), COORD)
    info = np.zeros(6, np.int32)
    hits = L.emu_patch_map_find(len(arrays), arrays.ctypes.data, len(params), params.ctypes.data, int(triangular), len(face),
                                face.ctypes.data, s.ctypes.data, t.ctypes.data, out.ctypes.data, info.ctypes.data)
    return hits, out, info


def assert_same_coords(got, want, s, t, what):
    miss_g, miss_w = got["arrayIndex"] < 0, want["arrayIndex"] < 0
    assert np.array_equal(miss_g, miss_w), f"{what}: hole / out-of-range samples differ"
    hit = ~miss_w
    for f in ("arrayIndex", "patchIndex", "vertIndex"):
        bad = np.nonzero(got[f][hit] != want[f][hit])[0]
        assert bad.size == 0, f"{what}: {f} differs at {bad[:5]} (s,t)={s[hit][bad[:5]]},{t[hit][bad[:5]]}"
    assert np.array_equal(got["s"][hit].view(np.int32), want["s"][hit].view(np.int32)), what
    assert np.array_equal(got["t"][hit].view(np.int32), want["t"][hit].view(np.int32)), what


@pytest.mark.parametrize("name", golden_names("patchmap_"))
def test_oracle_find_patches_vs_golden(name):
    """The C oracle's reference-style restatement (double-precision halving) against Far::PatchMap's recorded answers."""
    from oracle import oracle
    d = golden(name)
    got = oracle.find_patches(d["arrays"], d["params"], bool(d["triangular"]), d["face"], d["s"], d["t"])
    assert_same_coords(got, d["coords"], d["s"], d["t"], f"oracle {name}")


@pytest.mark.parametrize("name", golden_names("patchmap_"))
def test_patch_map_descent_vs_golden(name):
    L = _lib()
    d = golden(name)
    hits, got, info = emu_find(L, d["arrays"], d["params"], int(d["triangular"]), d["face"], d["s"], d["t"])
    assert hits == int((d["coords"]["arrayIndex"] >= 0).sum())
    assert_same_coords(got, d["coords"], d["s"], d["t"], name)
    assert info[5] == len(d["params"]) and info[3] == int(d["triangular"])
    assert 0 <= info[0] and info[1] < int(d["num_ptex_faces"])


def test_patch_map_rejects_bad_tables():
    L = _lib()
    d = golden("patchmap_catmark_cube")
    arrays = d["arrays"].copy()
    arrays["primitiveIdBase"][0] = 1                      # arrays no longer tile the parameter table
    rc, _, _ = emu_find(L, arrays, d["params"], 0, d["face"][:4], d["s"][:4], d["t"][:4])
    assert rc == -1
    params = np.concatenate([d["params"], d["params"][:1]])          # the same cell claimed twice
    arrays = d["arrays"].copy()
    arrays["numPatches"][-1] += 1
    rc, _, _ = emu_find(L, arrays, params, 0, d["face"][:4], d["s"][:4], d["t"][:4])
    assert rc == -2
    hits, out, info = emu_find(L, d["arrays"][:0], d["params"][:0], 0, d["face"][:4], d["s"][:4], d["t"][:4])
    assert hits == 0 and (out["arrayIndex"] == -1).all()              # empty table: every sample misses


REF_SHAPES = [("catmark_car", 3, "gregory"), ("catmark_helmet", 2, "bspline"), ("catmark_hole_test1", 3, "gregory"),
              ("catmark_hole_test3", 2, "gregory"), ("catmark_hole_test4", 3, "gregory"), ("catmark_pawn", 3, "gregory"),
              ("catmark_smoothtris0", 4, "gregory"), ("catmark_pyramid_creases1", 5, "gregory"),
              ("catmark_single_crease", 6, "gregory"), ("catmark_gregory_test7", 10, "gregory"),
              ("loop_pole64", 3, "gregory"), ("loop_toroidal_tet", 6, "gregory"), ("loop_saddle_edgecorner", 8, "gregory"),
              ("bilinear_cube", 3, "bilinear")]


@pytest.mark.parametrize("shape,level,endcap", REF_SHAPES)
def test_patch_map_descent_vs_reference(shape, level, endcap):
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference sources not present (GPU box)")
    from oracle import ref
    from tests.golden.make_golden import patch_map_samples
    L = _lib()
    m = ref.Mesh.from_shape(shape)
    for single_crease in (False, True):
        pt = m.patch_table(level, end_cap=endcap, fvar=False, inf_sharp=True, single_crease=single_crease,
                           legacy_sharp_corner=False, refine_first=True)
        face, s, t = patch_map_samples(m, 20000, 17 + level)
        want = m.find_patches(pt, face, s, t)
        hits, got, _ = emu_find(L, pt.vertex.arrays, pt.vertex.params, m.reg_face_size == 3, face, s, t)
        assert_same_coords(got, want, s, t, f"{shape} L{level} sc={single_crease}")
        assert hits == int((want["arrayIndex"] >= 0).sum())
        from oracle import oracle                         # pins the oracle's restatement against the live reference too
        assert_same_coords(oracle.find_patches(pt.vertex.arrays, pt.vertex.params, m.reg_face_size == 3, face, s, t), want, s, t,
                           f"oracle {shape} L{level} sc={single_crease}")
        m = ref.Mesh.from_shape(shape)                    # a refiner can be refined only once
