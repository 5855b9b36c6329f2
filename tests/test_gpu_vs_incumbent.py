"""A second, independent checker on the GPU box: the reference's OWN CUDA backend kernels (osd/cudaKernel.cu compiled
unmodified for sm_100a, oracle/_ref/libosdcudaref.so) against this library on the same device buffers."""
import numpy as np
import pytest
import torch

import opensubdiv_b200 as osd
from oracle import cuda_ref
from tests.gpu_util import D, dev, coords_dev, oracle_patches, oracle_stencils
from tests.util import golden, golden_names, table_from, triple_from, assert_close

pytestmark = pytest.mark.gpu
OUT6 = ("p", "du", "dv", "duu", "duv", "dvv")


def _need():
    if not cuda_ref.available():
        pytest.skip("oracle/_ref/libosdcudaref.so not built (needs /root/reference at build time)")


class _PT:
    def __init__(self, vertex):
        self.vertex, self.varying, self.fvar = vertex, None, []


@pytest.mark.parametrize("name", golden_names("stencils_"))
def test_stencils_match_reference_cuda_backend(name):
    _need()
    d = golden(name)
    t = table_from(d, "t_")
    src = np.ascontiguousarray(d["src"], np.float32)
    L = src.shape[1]
    ncv, n = t.num_control_verts, t.num_stencils
    tbl = osd.B200StencilTable.Create(t)
    ours = torch.zeros((ncv + n, L), device="cuda")
    ours[:ncv] = dev(src)
    theirs = ours.clone()
    assert osd.B200Evaluator.EvalStencils(ours, D(0, L, L), ours, D(ncv * L, L, L), tbl)
    sizes, offsets, indices, weights = (dev(x) for x in (t.sizes, t.offsets, t.indices, t.weights))
    torch.cuda.synchronize()
    # Osd::CudaEvaluator::EvalStencils (osd/cudaEvaluator.cpp:150-170): L = 3 / 4 packed take the tuned kernels
    cuda_ref.eval_stencils(theirs.data_ptr(), theirs.data_ptr() + ncv * L * 4, L, L, L, sizes.data_ptr(), offsets.data_ptr(),
                           indices.data_ptr(), weights.data_ptr(), 0, n)
    torch.cuda.synchronize()
    scale = oracle_stencils(src.reshape(-1), (0, L, L), n, L, t, 1, abs_scale=True)[0]
    assert_close(ours[ncv:].cpu().numpy(), theirs[ncv:].cpu().numpy(), scale, f"{name} vs reference CUDA kernel")
    assert_close(theirs[ncv:].cpu().numpy(), d["out"], scale, f"{name} reference CUDA kernel vs CpuEvaluator")


@pytest.mark.parametrize("name", golden_names("patches_"))
def test_patches_match_reference_cuda_backend(name):
    _need()
    d = golden(name)
    vtx = triple_from(d, "vtx_")
    coords = d["coords"]
    n = len(coords)
    pt = osd.B200PatchTable.Create(_PT(vtx))
    pc = coords_dev(coords)
    src = dev(d["vb"])
    ours = torch.zeros((n, 18), device="cuda")
    theirs = torch.zeros((n, 18), device="cuda")
    args = []
    for k in range(6):
        args += [ours, D(3 * k, 3, 18)]
    assert osd.B200Evaluator.EvalPatches(src, D(0, 3, 3), *args, n, pc, pt, None)
    torch.cuda.synchronize()
    cuda_ref.eval_patches(src.data_ptr(), [theirs.data_ptr() + 12 * k for k in range(6)], 3, 3, [18] * 6, n, pc.data_ptr(),
                          pt.GetPatchArrayBuffer(), pt.GetPatchIndexBuffer(), pt.GetPatchParamBuffer())
    torch.cuda.synchronize()
    scales = oracle_patches(d["vb"], (0, 3, 3), 3, coords, vtx, 6, abs_scale=True)
    a, b = ours.cpu().numpy(), theirs.cpu().numpy()
    for k in range(6):
        # the reference CUDA kernel itself differs from CpuEvaluator by fp contraction: both sides get the 1e-6 budget
        assert_close(a[:, 3 * k:3 * k + 3], b[:, 3 * k:3 * k + 3], scales[k], f"{name} {OUT6[k]} vs reference CUDA kernel", tol=2e-6)
        assert_close(b[:, 3 * k:3 * k + 3], d["out_" + OUT6[k]], scales[k], f"{name} {OUT6[k]} reference CUDA kernel vs CpuEvaluator")
