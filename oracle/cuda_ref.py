"""TEST / BENCH INFRASTRUCTURE.  ctypes access to the reference's own CUDA backend launchers (osd/cudaKernel.cu,
compiled unmodified for sm_100a into oracle/_ref/libosdcudaref.so by `make -C oracle/ref cuda`): the incumbent GPU
implementation.  Used as a second checker in tests/ and timed next to the product in bench.py; never on a product path.

Pointer arguments are raw device addresses (ints), already offset the way Osd::CudaEvaluator passes them
(osd/cudaEvaluator.cpp:150-292: src + srcDesc.offset, dst + dstDesc.offset)."""
import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libosdcudaref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, i = C.c_void_p, C.c_int
        L.CudaEvalStencils.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, i, i]
        L.CudaEvalStencils.restype = None
        L.CudaEvalPatches.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp]
        L.CudaEvalPatches.restype = None
        L.CudaEvalPatchesWithDerivatives.argtypes = [vp] * 7 + [i] * 8 + [i, vp, vp, vp, vp]
        L.CudaEvalPatchesWithDerivatives.restype = None
        _lib = L
    return _lib


def eval_stencils(src, dst, length, src_stride, dst_stride, sizes, offsets, indices, weights, start, end):
    """osd/cudaKernel.cu:351-383 (legacy default stream)."""
    lib().CudaEvalStencils(src, dst, length, src_stride, dst_stride, sizes, offsets, indices, weights, start, end)


def eval_patches(src, dsts, length, src_stride, dst_strides, n, coords, arrays, indices, params):
    """dsts / dst_strides: 1 entry -> CudaEvalPatches, else 6 entries (None = skipped) -> ...WithDerivatives
    (osd/cudaKernel.cu:387-424)."""
    if len(dsts) == 1:
        lib().CudaEvalPatches(src, dsts[0], length, src_stride, dst_strides[0], n, coords, arrays, indices, params)
        return
    d = list(dsts) + [None] * (6 - len(dsts))
    st = list(dst_strides) + [0] * (6 - len(dst_strides))
    lib().CudaEvalPatchesWithDerivatives(src, d[0], d[1], d[2], d[3], d[4], d[5], length, src_stride, st[0], st[1], st[2],
                                         st[3], st[4], st[5], n, coords, arrays, indices, params)
