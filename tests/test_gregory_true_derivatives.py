"""The reference's build option OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES (osd/patchBasis.h:421-487, far/patchBasis.cpp:462):
golden outputs of the reference compiled with it (tests/golden/truederiv_*.npz, made by
tests/golden/make_golden_true_derivatives.py from oracle/_ref/libosdref_td.so) against the oracle's restatement and the
host emulation of the CUDA kernel's arithmetic.  The GPU twin is tests/test_gpu_true_derivatives.py."""
import numpy as np
import pytest

from oracle import oracle, ref
from tests.util import golden, triple_from, assert_close
from tests.test_kernel_math_emu import _lib as emu_lib, OUT6

PATCH_SHAPES = ("catmark_cube_creases0", "catmark_gregory_test2", "catmark_car")
LIMIT_SHAPES = ("catmark_gregory_test2", "catmark_cube_creases0")
STREAMS = ("weights", "du", "dv", "duu", "duv", "dvv")


def oracle_eval(d, tr, nw, true_derivatives, scale=False):
    coords, vb = d["coords"], d["vb"]
    outs = [np.zeros((len(coords), 3), np.float32) for _ in range(nw)]

    def run():
        assert oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in outs], [(0, 3, 3)] * nw, coords,
                                   tr.arrays, tr.indices, tr.params)
    if true_derivatives:
        with oracle.gregory_true_derivatives():
            if scale:
                with oracle.abs_mode(2):
                    run()
            else:
                run()
    else:
        run()
    return outs


@pytest.mark.parametrize("shape", PATCH_SHAPES)
def test_oracle_is_bit_identical_to_the_reference_built_with_the_switch(shape):
    d, g = golden("patches_" + shape), golden("truederiv_" + shape)
    tr = triple_from(d, "vtx_")
    outs = oracle_eval(d, tr, 6, True)
    for k in range(6):
        assert np.array_equal(outs[k].view(np.int32), g["out_" + OUT6[k]].view(np.int32)), (shape, OUT6[k])
    # and the switch is off by default: the plain fixtures still come out
    outs = oracle_eval(d, tr, 6, False)
    for k in range(6):
        assert np.array_equal(outs[k].view(np.int32), d["out_" + OUT6[k]].view(np.int32)), (shape, OUT6[k])


@pytest.mark.parametrize("shape", LIMIT_SHAPES)
def test_oracle_limit_table_with_the_switch(shape):
    """Far::LimitStencilTableFactory of the switched build (far/patchBasis.cpp:483-530 feeds its weights)."""
    d, g = golden("limit_" + shape), golden("truederiv_limit_" + shape)
    if not ref.available():
        pytest.skip("needs oracle/_ref/libosdref.so to rebuild the patch table of the limit fixture")
    m = ref.Mesh.from_shape(shape).refine_adaptive(3)
    pt = m.patch_table(3, end_cap="gregory", refine_first=False)
    st = m.stencil_table(intermediate_levels=True, patch_table=pt)
    tri = m.reg_face_size == 3
    with oracle.gregory_true_derivatives():
        sizes, offsets, indices, ws = oracle.limit_stencil_table(
            pt.vertex.arrays, pt.vertex.indices, pt.vertex.params, tri, st.num_control_verts, st.sizes, st.offsets,
            st.indices, st.weights, d["face"], d["s"], d["t"], 6)
    assert np.array_equal(sizes, g["t_sizes"]) and np.array_equal(indices, g["t_indices"])
    for k, name in enumerate(("weights", "du", "dv", "duu", "duv", "dvv")):
        assert np.array_equal(ws[k].view(np.int32), g["t_" + name].view(np.int32)), (shape, name)
    assert not np.array_equal(g["t_du"], d["t_du"])            # the switch matters on this shape


@pytest.mark.parametrize("shape", PATCH_SHAPES)
@pytest.mark.parametrize("nw", [3, 6])
def test_kernel_math_with_the_switch(shape, nw):
    """The CUDA kernel's arithmetic (host emulation, tests/emu) with PatchIO::options = 1 against the switched reference,
    1e-6 relative in the conditioned scale of the switched weights (tests/util.py)."""
    import ctypes as C
    L = emu_lib()
    d, g = golden("patches_" + shape), golden("truederiv_" + shape)
    tr = triple_from(d, "vtx_")
    coords, vb = np.ascontiguousarray(d["coords"]), np.ascontiguousarray(d["vb"])
    outs = [np.zeros((len(coords), 3), np.float32) for _ in range(nw)]
    sd = (C.c_int * 3)(0, 3, 3)
    dd = (C.c_int * (3 * nw))(*([0, 3, 3] * nw))
    dptr = (C.c_void_p * nw)(*[o.ctypes.data for o in outs])
    a, ix, pr = (np.ascontiguousarray(x) for x in (tr.arrays, tr.indices, tr.params))
    L.emu_eval_patches_ex.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int]
    assert L.emu_eval_patches_ex(vb.ctypes.data, sd, nw, dptr, dd, len(coords), coords.ctypes.data, a.ctypes.data,
                                 ix.ctypes.data, pr.ctypes.data, 1) == 0
    scl = oracle_eval(d, tr, nw, True, scale=True)
    for k in range(nw):
        assert_close(outs[k], g["out_" + OUT6[k]], scl[k], f"{shape} {OUT6[k]} (true derivatives)")
    changed = max(float(np.abs(outs[k] - d["out_" + OUT6[k]]).max()) for k in range(1, nw))
    assert changed > 1e-3                                        # the option reached the kernel code


@pytest.mark.skipif(not ref.true_derivatives_available(), reason="oracle/_ref/libosdref_td.so not built")
def test_basis_weights_against_the_switched_reference_live():
    rng = np.random.default_rng(3)
    with ref.true_derivatives(), oracle.gregory_true_derivatives():
        for it in range(3000):
            a, b = rng.random(2).astype(np.float32)
            if it % 10 == 0:
                a = np.float32(rng.integers(0, 2))
            if it % 15 == 0:
                b = np.float32(rng.integers(0, 2))
            depth = int(rng.integers(0, 6))
            pu, pv = int(rng.integers(0, 1 << depth)), int(rng.integers(0, 1 << depth))
            f1 = depth | (pv << 12) | (pu << 22)
            frac = np.float32(1.0 / (1 << depth))
            s, t = float((pu + a) * frac), float((pv + b) * frac)
            n0, w0 = ref.osd_patch_basis(9, 0, f1, s, t)
            n1, w1 = oracle.patch_basis(9, 0, f1, s, t)
            assert n0 == n1 == 20
            for k in range(6):
                assert np.array_equal(w0[k].view(np.int32), w1[k].view(np.int32)), (it, k)
