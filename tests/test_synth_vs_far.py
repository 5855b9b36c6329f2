"""The synthetic table generators (opensubdiv_b200/synth.py, used by bench.py on the GPU box) against real Far tables."""
import numpy as np
import pytest

from oracle import oracle, ref
from opensubdiv_b200 import synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libosdref.so not built (needs /root/reference)")


def _canon(t):
    rowid = np.repeat(np.arange(len(t.sizes)), t.sizes)
    order = np.lexsort((t.indices, rowid))
    return t.indices[order], t.weights[order]


@pytest.mark.parametrize("scheme,nu,nv,level", [("catmark", 12, 10, 1), ("catmark", 9, 7, 3), ("loop", 8, 6, 2), ("loop", 12, 10, 3)])
def test_uniform_stencil_tables_equal_far(scheme, nu, nv, level):
    mesh = synth.torus_quads(nu, nv) if scheme == "catmark" else synth.torus_tris(nu, nv)
    mine = synth.uniform_stencil_table(mesh, level)
    m = ref.Mesh.from_topology(scheme, mesh.num_verts, np.full(len(mesh.faces), mesh.faces.shape[1], np.int32),
                               mesh.faces.reshape(-1))
    far = m.refine_uniform(level).stencil_table()
    assert np.array_equal(mine.sizes, far.sizes) and np.array_equal(mine.offsets, far.offsets)
    (i1, w1), (i2, w2) = _canon(mine), _canon(far)
    assert np.array_equal(i1, i2)
    assert np.abs(w1 - w2).max() <= 1e-7


def test_torus_patch_table_and_limit_stencils_equal_far():
    nu, nv = 7, 6
    mesh = synth.torus_quads(nu, nv)
    m = ref.Mesh.from_topology("catmark", mesh.num_verts, np.full(len(mesh.faces), 4, np.int32), mesh.faces.reshape(-1))
    pt = m.patch_table(3, end_cap="gregory")
    mine = synth.torus_patch_table(mesh)
    assert len(pt.vertex.arrays) == 1 and pt.vertex.arrays[0]["desc"] == 6
    assert np.array_equal(pt.vertex.indices, mine.vertex.indices)
    assert np.array_equal(pt.vertex.params["field1"], mine.vertex.params["field1"])
    assert np.array_equal(pt.vertex.params["field0"] & 0xFFFFFFF, mine.vertex.params["field0"])
    rng = np.random.default_rng(4)
    k = 400
    face = np.sort(rng.integers(0, nu * nv, k)).astype(np.int32)
    s, t = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    far = m.limit_stencil_table(face, s, t, True, True, patch_table=pt)
    mine_ls = synth.torus_limit_stencil_table(mesh, face, s, t)
    pos = mesh.positions
    for tbl_a, tbl_b in ((far, mine_ls),):
        a = [np.zeros((k, 3), np.float32) for _ in range(6)]
        b = [np.zeros((k, 3), np.float32) for _ in range(6)]
        oracle.eval_stencils(pos.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in a], [(0, 3, 3)] * 6, tbl_a.sizes,
                             tbl_a.offsets, tbl_a.indices, tbl_a.weight_streams(6))
        oracle.eval_stencils(pos.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in b], [(0, 3, 3)] * 6, tbl_b.sizes,
                             tbl_b.offsets, tbl_b.indices, tbl_b.weight_streams(6))
        for x, y in zip(a, b):
            assert np.abs(x - y).max() <= 2e-5 * max(1.0, np.abs(x).max())
