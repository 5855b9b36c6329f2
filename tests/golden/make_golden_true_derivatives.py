"""Golden outputs of the reference built with OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES (oracle/_ref/libosdref_td.so,
oracle/ref/Makefile), on the inputs of the committed patches_* / limit_* fixtures that contain GREGORY_BASIS patches:

    python tests/golden/make_golden_true_derivatives.py      (needs /root/reference -> make -C oracle/ref)

truederiv_<shape>.npz        Osd::CpuEvaluator::EvalPatches outputs (P, D1, D2) for patches_<shape>'s table and coords
truederiv_limit_<shape>.npz  Far::LimitStencilTableFactory table for limit_<shape>'s locations
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import ref                                   # noqa: E402
from tests.util import golden, triple_from, GOLDEN       # noqa: E402
from tests.golden.make_golden import table_dict          # noqa: E402

OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")
PATCH_SHAPES = ("catmark_cube_creases0", "catmark_gregory_test2", "catmark_car")
LIMIT_SHAPES = (("catmark_gregory_test2", 3), ("catmark_cube_creases0", 3))


def main():
    with ref.true_derivatives():
        for shape in PATCH_SHAPES:
            d = golden("patches_" + shape)
            tr = triple_from(d, "vtx_")
            assert (tr.arrays["desc"] == 9).any(), shape + " has no GREGORY_BASIS patches"
            coords, vb = d["coords"], d["vb"]
            outs = [np.zeros((len(coords), 3), np.float32) for _ in range(6)]
            assert ref.eval_patches(vb.reshape(-1), (0, 3, 3), [o.reshape(-1) for o in outs], [(0, 3, 3)] * 6, coords, tr)
            changed = max(float(np.abs(o - d["out_" + k]).max()) for k, o in zip(OUT6, outs))
            assert changed > 0, shape + ": the switch changed nothing"
            assert np.array_equal(outs[0], d["out_p"])          # positions do not depend on the switch
            np.savez_compressed(os.path.join(GOLDEN, "truederiv_" + shape + ".npz"), **{"out_" + k: o for k, o in zip(OUT6, outs)})
            print("truederiv_" + shape, "max change", changed)
        for shape, level in LIMIT_SHAPES:
            d = golden("limit_" + shape)
            m = ref.Mesh.from_shape(shape).refine_adaptive(level)
            st = m.limit_stencil_table(d["face"], d["s"], d["t"], first=True, second=True)
            np.savez_compressed(os.path.join(GOLDEN, "truederiv_limit_" + shape + ".npz"), **table_dict("t_", st))
            print("truederiv_limit_" + shape, st.num_stencils, "stencils; max |dw|", float(np.abs(st.du - d["t_du"]).max()))


if __name__ == "__main__":
    main()
