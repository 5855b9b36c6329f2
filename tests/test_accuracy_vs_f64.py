"""How far are the reference (fp32 Osd::CpuEvaluator, golden outputs) and the B200 patch-kernel arithmetic (host
emulation, tests/emu) from the SAME algorithm evaluated in double precision (oracle -DORACLE_F64)?

This substantiates the tolerance used by the parity tests: both implementations sit within ~1e-7 of the exact value
relative to the conditioned scale S, and the kernel is never meaningfully further from the truth than the reference."""
import numpy as np
import pytest

from oracle import oracle
from tests.test_kernel_math_emu import _lib, emu_patches
from tests.util import golden, golden_names, triple_from, table_from, weight_streams

OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")


def _rel(x, truth, scale):
    den = np.maximum(np.maximum(np.abs(truth), scale), 1e-30)
    return float((np.abs(x.astype(np.float64) - truth) / den).max())


@pytest.mark.parametrize("name", golden_names("patches_"))
def test_patch_kernel_is_as_close_to_double_precision_as_the_reference(name):
    L = _lib()
    d = golden(name)
    tr = triple_from(d, "vtx_")
    coords, vb = np.ascontiguousarray(d["coords"]), d["vb"]
    truth = oracle.eval_patches_f64(vb, (0, 3, 3), 3, coords, tr.arrays, tr.indices, tr.params, 6)
    got = emu_patches(L, vb.reshape(-1), (0, 3, 3), 3, coords, tr, 6)
    scale = [np.zeros((len(coords), 3), np.float32) for _ in range(6)]
    with oracle.abs_mode(2):
        oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in scale], [(0, 3, 3)] * 6, coords, tr.arrays,
                            tr.indices, tr.params)
    for k in range(6):
        e_ref = _rel(d["out_" + OUT6[k]], truth[k], scale[k])
        e_gpu = _rel(got[k], truth[k], scale[k])
        assert e_gpu <= 1e-6, (name, OUT6[k], e_gpu)
        assert e_ref <= 1e-6, (name, OUT6[k], e_ref)
        assert e_gpu <= 2.0 * e_ref + 5e-8, (name, OUT6[k], e_gpu, e_ref)      # never meaningfully worse than the reference


@pytest.mark.parametrize("name", golden_names("limit_"))
def test_reference_stencil_error_vs_double_precision(name):
    """The reference's own fp32 rounding error on derivative stencils, for the record (plain scale S = sum|w||x|)."""
    d = golden(name)
    t = table_from(d, "t_")
    n = t.num_stencils
    truth = oracle.eval_stencils_f64(d["src"], (0, 3, 3), n, 3, t.sizes, t.offsets, t.indices, weight_streams(t, 6))
    scale = [np.zeros((n, 3), np.float32) for _ in range(6)]
    with oracle.abs_mode(1):
        oracle.eval_stencils(d["src"].reshape(-1), (0, 3, 3), [o.reshape(-1) for o in scale], [(0, 3, 3)] * 6, t.sizes,
                             t.offsets, t.indices, weight_streams(t, 6))
    for k in range(6):
        assert _rel(d["out_" + OUT6[k]], truth[k], scale[k]) <= 1e-6


if __name__ == "__main__":       # prints the table quoted in DESIGN.md
    L = _lib()
    print("%-34s %s" % ("fixture", "  ".join("%-19s" % o for o in OUT6)))
    for name in golden_names("patches_"):
        d = golden(name)
        tr = triple_from(d, "vtx_")
        coords, vb = np.ascontiguousarray(d["coords"]), d["vb"]
        truth = oracle.eval_patches_f64(vb, (0, 3, 3), 3, coords, tr.arrays, tr.indices, tr.params, 6)
        got = emu_patches(L, vb.reshape(-1), (0, 3, 3), 3, coords, tr, 6)
        scale = [np.zeros((len(coords), 3), np.float32) for _ in range(6)]
        with oracle.abs_mode(2):
            oracle.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in scale], [(0, 3, 3)] * 6, coords, tr.arrays,
                                tr.indices, tr.params)
        cells = ["ref %.1e gpu %.1e" % (_rel(d["out_" + OUT6[k]], truth[k], scale[k]), _rel(got[k], truth[k], scale[k])) for k in range(6)]
        print("%-34s %s" % (name, "  ".join(cells)))
