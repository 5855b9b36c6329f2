"""Pins the oracle (and the product's patch-map code, host-emulated) against the reference compiled in place on RANDOM
topologies: open quad grids with randomly triangulated cells (non-quad faces -> ptex sub-faces), random semi-sharp and
infinitely sharp creases and corners, Catmark and Loop.  Stencils, EvalPatches and FindPatch must be bit-identical."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests.util import assert_close

if not os.path.isdir("/root/reference"):
    pytest.skip("needs the reference sources (compiled in place under oracle/_ref)", allow_module_level=True)
from oracle import ref  # noqa: E402


def random_mesh(seed, scheme):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(3, 7)), int(rng.integers(3, 7))
    vid = lambda i, j: j * (nx + 1) + i
    pos = np.zeros(((nx + 1) * (ny + 1), 3), np.float32)
    for j in range(ny + 1):
        for i in range(nx + 1):
            pos[vid(i, j)] = (i + 0.3 * rng.standard_normal(), j + 0.3 * rng.standard_normal(), rng.standard_normal())
    vpf, fv = [], []
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
            if scheme == "loop" or rng.random() < 0.2:          # split the cell (Catmark: triangles among the quads)
                if rng.random() < 0.5:
                    vpf += [3, 3]; fv += [a, b, c, a, c, d]
                else:
                    vpf += [3, 3]; fv += [a, b, d, b, c, d]
            else:
                vpf.append(4); fv += [a, b, c, d]
    # creases along random grid edges, corners on random vertices; sharpness 0.5 .. 4 or infinite (10)
    pairs, cw = [], []
    for _ in range(int(rng.integers(0, 6))):
        i, j = int(rng.integers(0, nx)), int(rng.integers(0, ny + 1))
        pairs += [vid(i, j), vid(i + 1, j)]
        cw.append(10.0 if rng.random() < 0.3 else float(rng.uniform(0.5, 4.0)))
    cv = rng.choice(len(pos), size=int(rng.integers(0, 4)), replace=False)
    cow = [10.0 if rng.random() < 0.5 else float(rng.uniform(0.5, 3.0)) for _ in cv]
    m = ref.Mesh.from_topology(scheme, len(pos), np.array(vpf, np.int32), np.array(fv, np.int32),
                               boundary_interp=int(rng.integers(1, 3)),
                               crease_pairs=np.array(pairs, np.int32) if pairs else None,
                               crease_weights=np.array(cw, np.float32) if cw else None,
                               corner_verts=np.array(cv, np.int32) if len(cv) else None,
                               corner_weights=np.array(cow, np.float32) if len(cv) else None)
    return m, pos


@pytest.mark.parametrize("scheme", ["catmark", "loop"])
@pytest.mark.parametrize("seed", range(6))
def test_random_topology_stencils_bit_identical(seed, scheme):
    m, pos = random_mesh(seed, scheme)
    m.refine_uniform(2)
    st = m.stencil_table(intermediate_levels=True)
    n, L = st.num_stencils, 3
    a, b = np.zeros((n, L), np.float32), np.zeros((n, L), np.float32)
    assert ref.eval_stencils(pos.reshape(-1), (0, L, L), [a.reshape(-1)], [(0, L, L)], st)
    assert oracle.eval_stencils(pos.reshape(-1), (0, L, L), [b.reshape(-1)], [(0, L, L)], st.sizes, st.offsets, st.indices,
                                [st.weights])
    assert np.array_equal(a.view(np.int32), b.view(np.int32))


@pytest.mark.parametrize("scheme", ["catmark", "loop"])
@pytest.mark.parametrize("seed", range(6))
def test_random_topology_patches_and_patch_map(seed, scheme):
    from tests.golden.make_golden import patch_map_samples
    from tests.test_kernel_math_emu import _lib
    from tests.test_patch_map import assert_same_coords, emu_find
    m, pos = random_mesh(100 + seed, scheme)
    level = 2 + seed % 3
    pt = m.patch_table(level, end_cap=("gregory", "bspline")[seed % 2] if scheme == "catmark" else "gregory",
                       inf_sharp=bool(seed % 2), single_crease=(seed % 3 == 0), legacy_sharp_corner=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    ncv, n = st.num_control_verts, st.num_stencils
    vb = np.zeros((ncv + n, 3), np.float32)
    vb[:ncv] = pos
    assert ref.eval_stencils(vb.reshape(-1), (0, 3, 3), [vb.reshape(-1)], [(ncv * 3, 3, 3)], st)
    face, s, t = patch_map_samples(m, 4000, seed)
    want = m.find_patches(pt, face, s, t)
    tri = m.reg_face_size == 3
    assert_same_coords(oracle.find_patches(pt.vertex.arrays, pt.vertex.params, tri, face, s, t), want, s, t, "oracle FindPatch")
    hits, got, _ = emu_find(_lib(), pt.vertex.arrays, pt.vertex.params, tri, face, s, t)
    assert_same_coords(got, want, s, t, "product patch map (host-emulated)")
    pc = np.ascontiguousarray(want[want["arrayIndex"] >= 0])
    assert len(pc) > 1000
    a = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    b = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    sc = [np.zeros((len(pc), 3), np.float32) for _ in range(6)]
    assert ref.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in a], [(0, 3, 3)] * 6, pc, pt.vertex)
    args = (pc, pt.vertex.arrays, pt.vertex.indices, pt.vertex.params)
    assert oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in b], [(0, 3, 3)] * 6, *args)
    with oracle.abs_mode(2):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in sc], [(0, 3, 3)] * 6, *args)
    for k, (x, y, z) in enumerate(zip(a, b, sc)):
        assert_close(y, x, z, f"{scheme} seed {seed} output {k}", tol=3e-7)      # Gregory triangle: 3e-7; everything else 0
    # the CUDA patch kernel's own arithmetic (host-emulated) on the same random tables, at the product's tolerance
    from tests.test_kernel_math_emu import emu_patches
    got6 = emu_patches(_lib(), vb.reshape(-1), (0, 3, 3), 3, pc, pt.vertex, 6)
    for k, (x, y, z) in enumerate(zip(a, got6, sc)):
        assert_close(y, x, z, f"kernel math {scheme} seed {seed} output {k}")
