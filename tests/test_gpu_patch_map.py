"""B200PatchMap (device-side Far::PatchMap::FindPatch) on a real B200, through the C ABI: bit-exact against
the reference's recorded FindPatch results, and the FindPatches -> EvalPatches pipeline with holes."""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from tests.gpu_util import D, dev, oracle_patches
from tests.test_patch_map import assert_same_coords
from tests.util import golden, golden_names, triple_from, assert_close

pytestmark = pytest.mark.gpu
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")


class _PT:
    def __init__(self, vertex, varying=None, fvar=None):
        self.vertex, self.varying, self.fvar = vertex, varying, fvar or []


class _Tr:
    def __init__(self, arrays, params, indices=None):
        self.arrays, self.params, self.indices = arrays, params, indices


def find_on_device(pm, face, s, t, aos=False):
    n = len(face)
    out = torch.zeros(n * 5, dtype=torch.int32, device="cuda")
    found = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    if aos:        # {int face; float s; float t} records, like glEvalLimit's particle positions
        rec = np.zeros(n, dtype=[("f", "<i4"), ("s", "<f4"), ("t", "<f4")])
        rec["f"], rec["s"], rec["t"] = face, s, t
        buf = torch.from_numpy(rec.view(np.int32).copy()).cuda()
        base = buf.data_ptr()
        assert pm.FindPatches(n, base, base + 4, base + 8, out, found, strides=(3, 3, 3))
    else:
        assert pm.FindPatches(n, dev(face), dev(s), dev(t), out, found)
    torch.cuda.synchronize()
    return out.cpu().numpy().view(osd.PATCH_COORD_DTYPE), int(found.item()), out


@pytest.mark.parametrize("name", golden_names("patchmap_"))
def test_find_patches_vs_reference_golden(name):
    d = golden(name)
    pm = osd.B200PatchMap.Create(_PT(_Tr(d["arrays"], d["params"])), patchesAreTriangular=bool(d["triangular"]))
    want = d["coords"]
    for aos in (False, True):
        got, found, _ = find_on_device(pm, d["face"], d["s"], d["t"], aos)
        assert_same_coords(got, want, d["s"], d["t"], f"{name} aos={aos}")
        assert found == int((want["arrayIndex"] >= 0).sum())
        from oracle import oracle
        assert_same_coords(got, oracle.find_patches(d["arrays"], d["params"], bool(d["triangular"]), d["face"], d["s"], d["t"]),
                           d["s"], d["t"], f"{name} vs oracle")
        miss = want["arrayIndex"] < 0          # misses keep their (s, t)
        assert np.array_equal(got["s"][miss], d["s"][miss]) and np.array_equal(got["t"][miss], d["t"][miss])
    assert pm.GetNumPatches() == len(d["params"]) and pm.GetMaxDepth() == int((d["params"]["field1"] & 0xf).max())
    # ragged sizes around the warp / block granularity, and the empty call
    for n in (1, 31, 33, 255, 257):
        got, found, _ = find_on_device(pm, d["face"][:n], d["s"][:n], d["t"][:n])
        assert_same_coords(got, want[:n], d["s"][:n], d["t"][:n], f"{name} n={n}")
    assert pm.FindPatches(0, None, None, None, None)


@pytest.mark.parametrize("name", ["patches_catmark_car", "patches_loop_icosahedron", "patches_catmark_gregory_test2"])
def test_find_then_eval_pipeline_with_holes(name):
    """(face, s, t) -> FindPatches -> EvalPatches without leaving the device; samples that hit nothing keep their outputs."""
    d = golden(name)
    vtx = triple_from(d, "vtx_")
    var = triple_from(d, "var_") if "var_arrays" in d.files else None
    pt = osd.B200PatchTable.Create(_PT(vtx, var))
    pm = osd.B200PatchMap.Create(_PT(vtx, var))
    coords = d["coords"]
    n0 = len(coords)
    face = (vtx.params["field0"][coords["patchIndex"]] & 0x0fffffff).astype(np.int32)
    # every third sample is moved to a face the table does not have
    rng = np.random.default_rng(1)
    miss = np.zeros(n0, bool)
    miss[::3] = True
    face = np.where(miss, np.where(rng.random(n0) < 0.5, -1, face.max() + 1 + rng.integers(0, 5, n0)), face).astype(np.int32)
    got, found, pc = find_on_device(pm, face, coords["s"], coords["t"])
    assert found == int((~miss).sum())
    assert np.array_equal(got["arrayIndex"] < 0, miss)
    assert np.array_equal(got[~miss].view(np.int32), np.ascontiguousarray(coords[~miss]).view(np.int32))
    src = dev(d["vb"])
    scales = oracle_patches(d["vb"], (0, 3, 3), 3, coords, vtx, 6, abs_scale=True)
    for nw in (1, 3, 6):
        out = torch.full((n0, 3 * nw), -123.0, device="cuda")
        args = []
        for k in range(nw):
            args += [out, D(3 * k, 3, 3 * nw)]
        assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n0, pc, pt, None)
        res = out.cpu().numpy()
        assert (res[miss] == -123.0).all(), "outputs of samples outside every patch must stay untouched"
        for k in range(nw):
            assert_close(res[~miss, 3 * k:3 * k + 3], d["out_" + OUT6[k]][~miss], scales[k][~miss], f"{name} nw={nw} {OUT6[k]}")


def test_patch_map_create_rejects_inconsistent_table():
    d = golden("patchmap_catmark_cube")
    arrays = d["arrays"].copy()
    arrays["primitiveIdBase"][0] = 3
    with pytest.raises(osd.B200OsdError):
        osd.B200PatchMap.Create(_PT(_Tr(arrays, d["params"])), patchesAreTriangular=False)
