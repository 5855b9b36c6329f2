"""TEST INFRASTRUCTURE -- ctypes view of oracle/liboracle.so (oracle/osd_oracle.c).

The C oracle is a single-threaded restatement of the reference CPU algorithm (see the header of
osd_oracle.c for the file:line map).  It travels to the GPU box, where /root/reference is absent, and is
the checker for every `-m gpu` parity test.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; opensubdiv_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC_PATH = os.path.join(_HERE, "osd_oracle.c")
LIB_PATH = os.path.join(_HERE, "liboracle.so")
LIB64_PATH = os.path.join(_HERE, "liboracle_f64.so")

DESC_DTYPE = np.dtype([("offset", "<i4"), ("length", "<i4"), ("stride", "<i4")])

_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off: separate multiply/add like the reference's x86-64 build."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC_PATH):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99", "-Wall",
                               "-o", LIB_PATH, SRC_PATH, "-lm"])
    return LIB_PATH


_lib64 = None


def lib64():
    """The same restatement compiled with -DORACLE_F64 (all arithmetic in double on the same fp32 inputs): the
    "truth" used to measure the rounding error of the reference and of the B200 kernels.  Not the parity oracle."""
    global _lib64
    if _lib64 is None:
        if not os.path.exists(LIB64_PATH) or os.path.getmtime(LIB64_PATH) < os.path.getmtime(SRC_PATH):
            cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
            subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99", "-Wall", "-DORACLE_F64",
                                   "-o", LIB64_PATH, SRC_PATH, "-lm"])
        L = C.CDLL(LIB64_PATH)
        vp = C.c_void_p
        L.oracle_eval_stencils.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        L.oracle_eval_patches.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        L.oracle_set_abs_mode.argtypes = [C.c_int]
        L.oracle_set_gregory_true_derivatives.argtypes = [C.c_int]
        _lib64 = L
    return _lib64


def eval_stencils_f64(src, src_desc, n_rows, L, sizes, offsets, indices, weights, start=0, end=None):
    """Double-precision evaluation of the same table on the same fp32 inputs -> list of [n_rows, L] float64 arrays."""
    nw = len(weights)
    end = len(sizes) if end is None else end
    src64 = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
    w64 = [np.ascontiguousarray(w, dtype=np.float64) for w in weights]
    outs = [np.zeros((n_rows, L), np.float64) for _ in range(nw)]
    sd, dd = _descs([src_desc]), _descs([(0, L, L)] * nw)
    dptr = (C.c_void_p * nw)(*[o.ctypes.data for o in outs])
    wptr = (C.c_void_p * nw)(*[w.ctypes.data for w in w64])
    assert lib64().oracle_eval_stencils(nw, _p(src64), _p(sd), dptr, _p(dd), _p(sizes), _p(offsets), _p(indices), wptr, start, end)
    return outs


def eval_patches_f64(src, src_desc, L, coords, arrays, indices, params, nw=6):
    src64 = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
    outs = [np.zeros((len(coords), L), np.float64) for _ in range(nw)]
    sd, dd = _descs([src_desc]), _descs([(0, L, L)] * nw)
    dptr = (C.c_void_p * nw)(*[o.ctypes.data for o in outs])
    assert lib64().oracle_eval_patches(nw, _p(src64), _p(sd), dptr, _p(dd), len(coords), _p(coords), _p(arrays), _p(indices), _p(params))
    return outs


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.oracle_eval_stencils.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        L.oracle_eval_patches.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        L.oracle_patch_basis.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_float] + [vp] * 6
        L.oracle_version.restype = C.c_char_p
        L.oracle_set_abs_mode.argtypes = [C.c_int]
        L.oracle_set_gregory_true_derivatives.argtypes = [C.c_int]
        L.oracle_patch_map_create.restype = vp
        L.oracle_patch_map_create.argtypes = [C.c_int, vp, C.c_int, vp, C.c_int]
        L.oracle_patch_map_find.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.oracle_patch_map_free.argtypes = [vp]
        L.oracle_limit_table_create.restype = vp
        L.oracle_limit_table_create.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp]
        L.oracle_limit_table_free.argtypes = [vp]
        L.oracle_limit_table_num_stencils.argtypes = [vp]
        L.oracle_limit_table_num_elements.argtypes = [vp]
        L.oracle_limit_table_ints.restype = C.POINTER(C.c_int)
        L.oracle_limit_table_ints.argtypes = [vp, C.c_int]
        L.oracle_limit_table_weights.restype = C.POINTER(C.c_float)
        L.oracle_limit_table_weights.argtypes = [vp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _descs(descs) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(descs, dtype=np.int32).reshape(-1, 3))


def eval_stencils(src: np.ndarray, src_desc, dsts: Sequence[Optional[np.ndarray]], dst_descs,
                  sizes, offsets, indices, weights: Sequence[np.ndarray], start: int = 0,
                  end: Optional[int] = None) -> bool:
    """Restates Osd::CpuEvaluator::EvalStencils (osd/cpuEvaluator.cpp:37-125 -> osd/cpuKernel.cpp:71-240)."""
    nw = len(dsts)
    end = len(sizes) if end is None else end
    sd, dd = _descs([src_desc]), _descs(dst_descs)
    dptr = (C.c_void_p * nw)(*[None if d is None else d.ctypes.data for d in dsts])
    wptr = (C.c_void_p * nw)(*[w.ctypes.data for w in weights[:nw]])
    return bool(lib().oracle_eval_stencils(nw, _p(src), _p(sd), dptr, _p(dd), _p(sizes), _p(offsets), _p(indices),
                                           wptr, start, end))


def eval_patches(src: np.ndarray, src_desc, dsts: Sequence[Optional[np.ndarray]], dst_descs,
                 coords: np.ndarray, arrays: np.ndarray, indices: np.ndarray, params: np.ndarray) -> bool:
    """Restates Osd::CpuEvaluator::EvalPatches (osd/cpuEvaluator.cpp:157-381)."""
    nw = len(dsts)
    sd, dd = _descs([src_desc]), _descs(dst_descs)
    dptr = (C.c_void_p * nw)(*[None if d is None else d.ctypes.data for d in dsts])
    return bool(lib().oracle_eval_patches(nw, _p(src), _p(sd), dptr, _p(dd), len(coords), _p(coords), _p(arrays),
                                          _p(indices), _p(params)))


def patch_basis(patch_type: int, field0: int, field1: int, s: float, t: float, nw: int = 6):
    """Restates OsdEvaluatePatchBasis (osd/patchBasis.h:1555-1610)."""
    w = [np.zeros(20, dtype=np.float32) for _ in range(6)]
    ptrs = [_p(x) for x in w[:nw]] + [None] * (6 - nw)
    n = lib().oracle_patch_basis(patch_type, int(field0) & 0xFFFFFFFF, int(field1) & 0xFFFFFFFF, s, t, *ptrs)
    return n, w[:nw]


class gregory_true_derivatives:
    """Context manager: inside it the Gregory basis uses the reference's OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES form
    (osd/patchBasis.h:441-487) in every entry of the fp32 oracle (patch_basis, eval_patches, limit_stencil_table)."""

    def __enter__(self):
        lib().oracle_set_gregory_true_derivatives(1)
        lib64().oracle_set_gregory_true_derivatives(1)

    def __exit__(self, *exc):
        lib().oracle_set_gregory_true_derivatives(0)
        lib64().oracle_set_gregory_true_derivatives(0)


class abs_mode:
    """Context manager: inside it eval_stencils / eval_patches return the tolerance scale instead of the value.
    kind=1: S = sum_j |w_j||x_j|.  kind=2 (patches): S = sum_j (|w_j| + W)|x_j| with W = max |unfolded weight| of the
    set -- the weights themselves are formed with cancellation (boundary folding, B-spline polynomials near knots) so
    every correct fp32 evaluation, the reference's included, carries an error ~eps*W per weight."""

    def __init__(self, kind: int = 1):
        self.kind = kind

    def __enter__(self):
        lib().oracle_set_abs_mode(self.kind)

    def __exit__(self, *exc):
        lib().oracle_set_abs_mode(0)


COORD_DTYPE = np.dtype([("arrayIndex", "<i4"), ("patchIndex", "<i4"), ("vertIndex", "<i4"), ("s", "<f4"), ("t", "<f4")])


def find_patches(arrays: np.ndarray, params: np.ndarray, triangular: bool, ptex_face, s, t) -> np.ndarray:
    """Restates Far::PatchMap (far/patchMap.cpp:96-188) + FindPatch (far/patchMap.h:180-217) + Osd::PatchCoord
    (osd/types.h:53-54): one 20-byte record per sample, arrayIndex = -1 where the reference returns a NULL handle."""
    arrays, params = np.ascontiguousarray(arrays), np.ascontiguousarray(params)
    f = np.ascontiguousarray(ptex_face, dtype=np.int32)
    ss, tt = np.ascontiguousarray(s, dtype=np.float32), np.ascontiguousarray(t, dtype=np.float32)
    out = np.zeros(len(f), COORD_DTYPE)
    L = lib()
    h = L.oracle_patch_map_create(len(arrays), _p(arrays) if len(arrays) else None, len(params),
                                  _p(params) if len(params) else None, int(bool(triangular)))
    assert h, "oracle_patch_map_create failed"
    try:
        L.oracle_patch_map_find(h, len(f), _p(f), _p(ss), _p(tt), _p(out))
    finally:
        L.oracle_patch_map_free(h)
    return out


def limit_stencil_table(arrays, patch_indices, params, triangular, num_control_verts, cv_sizes, cv_offsets, cv_indices,
                        cv_weights, ptex_face, s, t, nw: int = 6):
    """Restates the per-location loop of Far::LimitStencilTableFactory::Create (far/stencilTableFactory.cpp:559-662)
    and StencilBuilder's merge (far/stencilBuilder.cpp): returns (sizes, offsets, indices, [weights x nw]) for the
    locations FindPatch resolves, in order.  cv_* = the refined + local-point stencil table WITHOUT control-vertex rows
    (row r belongs to vertex num_control_verts + r), feature-adaptive refinement."""
    arrays, params = np.ascontiguousarray(arrays), np.ascontiguousarray(params)
    pix = np.ascontiguousarray(patch_indices, dtype=np.int32)
    cs, co, ci = (np.ascontiguousarray(x, dtype=np.int32) for x in (cv_sizes, cv_offsets, cv_indices))
    cw = np.ascontiguousarray(cv_weights, dtype=np.float32)
    f = np.ascontiguousarray(ptex_face, dtype=np.int32)
    ss, tt = np.ascontiguousarray(s, dtype=np.float32), np.ascontiguousarray(t, dtype=np.float32)
    L = lib()
    pm = L.oracle_patch_map_create(len(arrays), _p(arrays), len(params), _p(params), int(bool(triangular)))
    assert pm
    h = None
    try:
        h = L.oracle_limit_table_create(nw, pm, _p(arrays), _p(pix), _p(params), int(num_control_verts), _p(cs), _p(co), _p(ci),
                                        _p(cw), len(f), _p(f), _p(ss), _p(tt))
        assert h, "oracle_limit_table_create failed"
        n, ne = L.oracle_limit_table_num_stencils(h), L.oracle_limit_table_num_elements(h)
        ints = [np.ctypeslib.as_array(L.oracle_limit_table_ints(h, k), shape=(m,)).copy() if m else np.zeros(0, np.int32)
                for k, m in ((0, n), (1, n), (2, ne))]
        ws = [np.ctypeslib.as_array(L.oracle_limit_table_weights(h, k), shape=(ne,)).copy() if ne else np.zeros(0, np.float32)
              for k in range(nw)]
        return ints[0], ints[1], ints[2], ws
    finally:
        if h:
            L.oracle_limit_table_free(h)
        L.oracle_patch_map_free(pm)
