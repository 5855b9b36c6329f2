"""TEST INFRASTRUCTURE -- ctypes view of oracle/_ref/libosdref.so.

libosdref.so is the UNMODIFIED OpenSubdiv 3.6.0 reference CPU path compiled in place from
/root/reference by oracle/ref/Makefile (plus oracle/ref/ref_shim.cpp).  It is used ONLY as
  * the producer of real Far tables for tests / golden fixtures, and
  * the ground-truth checker / CPU baseline (Osd::CpuEvaluator, Osd::OmpEvaluator).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (opensubdiv_b200/) never does.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libosdref.so")
# the same sources compiled with -DOPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES (oracle/ref/Makefile)
LIB_PATH_TD = os.path.join(_HERE, "_ref", "libosdref_td.so")

# numpy mirrors of the Osd POD types (osd/types.h:42-130, osd/patchBasisTypes.h:249-288)
PATCH_COORD_DTYPE = np.dtype([("arrayIndex", "<i4"), ("patchIndex", "<i4"), ("vertIndex", "<i4"),
                              ("s", "<f4"), ("t", "<f4")])
PATCH_ARRAY_DTYPE = np.dtype([("regDesc", "<i4"), ("desc", "<i4"), ("numPatches", "<i4"),
                              ("indexBase", "<i4"), ("stride", "<i4"), ("primitiveIdBase", "<i4")])
PATCH_PARAM_DTYPE = np.dtype([("field0", "<u4"), ("field1", "<u4"), ("sharpness", "<f4")])

_libs = {}
_active = LIB_PATH


def available() -> bool:
    return os.path.exists(LIB_PATH)


def true_derivatives_available() -> bool:
    return os.path.exists(LIB_PATH_TD)


@contextlib.contextmanager
def true_derivatives():
    """Inside this context every call goes to the reference built with OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES.
    Handles (Mesh, tables) stay with the library that made them: create and use them inside the same context."""
    global _active
    prev, _active = _active, LIB_PATH_TD
    try:
        yield
    finally:
        _active = prev


def lib():
    path = _active
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not built (make -C oracle/ref needs /root/reference)")
    L = C.CDLL(path)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.ref_shape_name.restype = C.c_char_p
    L.ref_mesh_from_shape.restype = vp
    L.ref_mesh_from_shape.argtypes = [C.c_char_p]
    L.ref_mesh_from_shape_tiled.restype = vp
    L.ref_mesh_from_shape_tiled.argtypes = [C.c_char_p, C.c_int]
    L.ref_mesh_from_topology.restype = vp
    L.ref_mesh_from_topology.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp,
                                         C.c_int, vp, vp, C.c_int, vp, vp]
    L.ref_mesh_free.argtypes = [vp]
    for n in ("num_base_verts", "num_base_faces", "reg_face_size", "num_fvar_channels", "num_verts_total",
              "max_level", "num_uvs", "num_ptex_faces"):
        getattr(L, "ref_mesh_" + n).argtypes = [vp]
    L.ref_mesh_num_base_fvar_values.argtypes = [vp, C.c_int]
    L.ref_mesh_level_num_verts.argtypes = [vp, C.c_int]
    L.ref_mesh_level_num_faces.argtypes = [vp, C.c_int]
    L.ref_mesh_level_face_verts.argtypes = [vp, C.c_int, vp]
    L.ref_mesh_positions.restype = fp
    L.ref_mesh_positions.argtypes = [vp]
    L.ref_mesh_uvs.restype = fp
    L.ref_mesh_uvs.argtypes = [vp]
    L.ref_mesh_refine_uniform.argtypes = [vp, C.c_int, C.c_int]
    L.ref_mesh_refine_adaptive.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ref_patch_table_create.restype = vp
    L.ref_patch_table_create.argtypes = [vp] + [C.c_int] * 8
    L.ref_patch_table_free.argtypes = [vp]
    for n in ("num_arrays", "num_indices", "num_params"):
        getattr(L, "ref_patch_table_" + n).argtypes = [vp, C.c_int]
    for n in ("arrays", "indices", "params"):
        f = getattr(L, "ref_patch_table_" + n)
        f.argtypes = [vp, C.c_int]
        f.restype = vp
    L.ref_patch_table_num_fvar_channels.argtypes = [vp]
    L.ref_patch_table_num_local_points.argtypes = [vp]
    L.ref_patch_table_num_local_points_fvar.argtypes = [vp, C.c_int]
    L.ref_patch_map_find.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.ref_patch_table_far_basis.argtypes = [vp, C.c_int, vp] + [vp] * 6
    L.ref_stencil_table_create.restype = vp
    L.ref_stencil_table_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.ref_limit_stencil_table_create.restype = vp
    L.ref_limit_stencil_table_create.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp]
    L.ref_stencil_table_free.argtypes = [vp]
    for n in ("num_stencils", "num_control_verts", "num_elements"):
        getattr(L, "ref_stencil_table_" + n).argtypes = [vp]
    for n in ("sizes", "offsets", "indices", "weights"):
        f = getattr(L, "ref_stencil_table_" + n)
        f.argtypes = [vp]
        f.restype = vp
    L.ref_stencil_table_deriv_weights.argtypes = [vp, C.c_int]
    L.ref_stencil_table_deriv_weights.restype = vp
    L.ref_stencil_table_update_values_xyz.argtypes = [vp, vp, vp]
    L.ref_eval_stencils.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
    L.ref_eval_patches.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.ref_osd_patch_basis.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float] + [vp] * 6
    L.ref_omp_set_threads.argtypes = [C.c_int]
    L.ref_version.restype = C.c_char_p
    _libs[path] = L
    return L


def _np_from(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    dt = np.dtype(dtype)
    buf = (C.c_char * (n * dt.itemsize)).from_address(ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value)
    return np.frombuffer(buf, dtype=dt, count=n).copy()


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def shape_names():
    L = lib()
    return [L.ref_shape_name(i).decode() for i in range(L.ref_num_shapes())]


# ------------------------------------------------------------------------------------ tables --
@dataclass
class StencilTable:
    """Flat copy of a Far::StencilTable / LimitStencilTable (far/stencilTable.h:156-186, 434-456)."""
    num_control_verts: int
    sizes: np.ndarray
    offsets: np.ndarray
    indices: np.ndarray
    weights: np.ndarray
    du: Optional[np.ndarray] = None
    dv: Optional[np.ndarray] = None
    duu: Optional[np.ndarray] = None
    duv: Optional[np.ndarray] = None
    dvv: Optional[np.ndarray] = None

    @property
    def num_stencils(self) -> int:
        return int(self.sizes.shape[0])

    def weight_streams(self, nw: int):
        return [self.weights, self.du, self.dv, self.duu, self.duv, self.dvv][:nw]


@dataclass
class PatchTriple:
    """(PatchArray[], index buffer, PatchParam[]) as Osd::CpuPatchTable exposes them (osd/cpuPatchTable.h)."""
    arrays: np.ndarray
    indices: np.ndarray
    params: np.ndarray


@dataclass
class PatchTable:
    vertex: PatchTriple
    varying: Optional[PatchTriple]
    fvar: list = field(default_factory=list)
    num_local_points: int = 0
    handle: object = None


class Mesh:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference could not create the mesh")
        self.h = handle
        self._L = lib()

    @classmethod
    def from_shape(cls, name: str) -> "Mesh":
        return cls(lib().ref_mesh_from_shape(name.encode()))

    @classmethod
    def from_shape_tiled(cls, name: str, copies: int) -> "Mesh":
        """`copies` replicas of a regression shape side by side (creases, corners and UVs replicated)."""
        return cls(lib().ref_mesh_from_shape_tiled(name.encode(), copies))

    @classmethod
    def from_topology(cls, scheme: str, num_verts: int, verts_per_face: np.ndarray, face_verts: np.ndarray,
                      boundary_interp: int = 1, fvar_linear_interp: int = 1,
                      fvar_indices: Optional[np.ndarray] = None, num_fvar_values: int = 0,
                      crease_pairs: Optional[np.ndarray] = None, crease_weights: Optional[np.ndarray] = None,
                      corner_verts: Optional[np.ndarray] = None, corner_weights: Optional[np.ndarray] = None):
        sc = {"bilinear": 0, "catmark": 1, "loop": 2}[scheme]
        vpf = np.ascontiguousarray(verts_per_face, dtype=np.int32)
        fv = np.ascontiguousarray(face_verts, dtype=np.int32)
        fvi = None if fvar_indices is None else np.ascontiguousarray(fvar_indices, dtype=np.int32)
        cp = None if crease_pairs is None else np.ascontiguousarray(crease_pairs, dtype=np.int32)
        cw = None if crease_weights is None else np.ascontiguousarray(crease_weights, dtype=np.float32)
        cv = None if corner_verts is None else np.ascontiguousarray(corner_verts, dtype=np.int32)
        cow = None if corner_weights is None else np.ascontiguousarray(corner_weights, dtype=np.float32)
        h = lib().ref_mesh_from_topology(sc, num_verts, len(vpf), _p(vpf), _p(fv), boundary_interp,
                                         fvar_linear_interp, num_fvar_values, _p(fvi),
                                         0 if cw is None else len(cw), _p(cp), _p(cw),
                                         0 if cow is None else len(cow), _p(cv), _p(cow))
        m = cls(h)
        m._keep = (vpf, fv, fvi, cp, cw, cv, cow)
        return m

    def __del__(self):
        try:
            if self.h:
                self._L.ref_mesh_free(self.h)
                self.h = None
        except Exception:
            pass

    # -- queries
    @property
    def num_base_verts(self):
        return self._L.ref_mesh_num_base_verts(self.h)

    @property
    def reg_face_size(self):
        return self._L.ref_mesh_reg_face_size(self.h)

    @property
    def num_verts_total(self):
        return self._L.ref_mesh_num_verts_total(self.h)

    @property
    def max_level(self):
        return self._L.ref_mesh_max_level(self.h)

    def level_num_verts(self, level):
        return self._L.ref_mesh_level_num_verts(self.h, level)

    def level_face_verts(self, level):
        n = self._L.ref_mesh_level_face_verts(self.h, level, None)
        out = np.zeros(n, dtype=np.int32)
        self._L.ref_mesh_level_face_verts(self.h, level, _p(out))
        return out

    def num_base_fvar_values(self, ch=0):
        return self._L.ref_mesh_num_base_fvar_values(self.h, ch)

    @property
    def num_fvar_channels(self):
        return self._L.ref_mesh_num_fvar_channels(self.h)

    @property
    def positions(self) -> np.ndarray:
        return _np_from(self._L.ref_mesh_positions(self.h), self.num_base_verts * 3, np.float32).reshape(-1, 3)

    @property
    def uvs(self) -> np.ndarray:
        return _np_from(self._L.ref_mesh_uvs(self.h), self._L.ref_mesh_num_uvs(self.h) * 2, np.float32).reshape(-1, 2)

    @property
    def num_ptex_faces(self):
        return self._L.ref_mesh_num_ptex_faces(self.h)

    # -- refinement
    def refine_uniform(self, level: int, full_topology_in_last_level: bool = False):
        self._L.ref_mesh_refine_uniform(self.h, level, int(full_topology_in_last_level))
        return self

    def refine_adaptive(self, level: int, single_crease=False, inf_sharp=False, consider_fvar=False):
        self._L.ref_mesh_refine_adaptive(self.h, level, int(single_crease), int(inf_sharp), int(consider_fvar))
        return self

    # -- tables
    def _stencils_from_handle(self, h, limit=False) -> StencilTable:
        L = self._L
        if not h:
            raise RuntimeError("reference returned no stencil table")
        n = L.ref_stencil_table_num_stencils(h)
        ne = L.ref_stencil_table_num_elements(h)
        st = StencilTable(
            num_control_verts=L.ref_stencil_table_num_control_verts(h),
            sizes=_np_from(L.ref_stencil_table_sizes(h), n, np.int32),
            offsets=_np_from(L.ref_stencil_table_offsets(h), n, np.int32),
            indices=_np_from(L.ref_stencil_table_indices(h), ne, np.int32),
            weights=_np_from(L.ref_stencil_table_weights(h), ne, np.float32))
        if limit:
            for k, name in enumerate(("du", "dv", "duu", "duv", "dvv"), start=1):
                p = L.ref_stencil_table_deriv_weights(h, k)
                setattr(st, name, _np_from(p, ne, np.float32) if p else None)
        L.ref_stencil_table_free(h)
        return st

    def stencil_table(self, mode: str = "vertex", intermediate_levels: bool = False, factorize: bool = True,
                      fvar_channel: int = 0, patch_table: Optional[PatchTable] = None) -> StencilTable:
        md = {"vertex": 0, "varying": 1, "fvar": 2}[mode]
        ph = patch_table.handle if patch_table is not None else None
        h = self._L.ref_stencil_table_create(self.h, md, int(intermediate_levels), int(factorize), fvar_channel, ph)
        return self._stencils_from_handle(h)

    def limit_stencil_table(self, ptex_face, s, t, first=True, second=False,
                            patch_table: Optional[PatchTable] = None) -> StencilTable:
        pf = np.ascontiguousarray(ptex_face, dtype=np.int32)
        ss = np.ascontiguousarray(s, dtype=np.float32)
        tt = np.ascontiguousarray(t, dtype=np.float32)
        ph = patch_table.handle if patch_table is not None else None
        h = self._L.ref_limit_stencil_table_create(self.h, len(pf), _p(pf), _p(ss), _p(tt), int(first), int(second), ph)
        return self._stencils_from_handle(h, limit=True)

    def patch_table(self, level: int, end_cap: str = "gregory", fvar: bool = False, fvar_legacy_linear: bool = True,
                    inf_sharp: bool = False, single_crease: bool = False, legacy_sharp_corner: bool = True,
                    refine_first: bool = True) -> PatchTable:
        ec = {"none": 0, "bilinear": 1, "bspline": 2, "gregory": 3, "legacy_gregory": 4}[end_cap]
        L = self._L
        h = L.ref_patch_table_create(self.h, level, ec, int(fvar), int(fvar_legacy_linear), int(inf_sharp),
                                     int(single_crease), int(legacy_sharp_corner), int(refine_first))
        if not h:
            raise RuntimeError("reference returned no patch table")

        def triple(which):
            na = L.ref_patch_table_num_arrays(h, which)
            pa = L.ref_patch_table_arrays(h, which)
            if not pa:
                return None
            return PatchTriple(
                arrays=_np_from(pa, na, PATCH_ARRAY_DTYPE),
                indices=_np_from(L.ref_patch_table_indices(h, which), L.ref_patch_table_num_indices(h, which), np.int32),
                params=_np_from(L.ref_patch_table_params(h, which), L.ref_patch_table_num_params(h, which), PATCH_PARAM_DTYPE))

        pt = PatchTable(vertex=triple(0), varying=triple(1),
                        fvar=[triple(2 + c) for c in range(L.ref_patch_table_num_fvar_channels(h))],
                        num_local_points=L.ref_patch_table_num_local_points(h), handle=h)
        return pt

    def find_patches(self, pt: PatchTable, ptex_face, s, t) -> np.ndarray:
        pf = np.ascontiguousarray(ptex_face, dtype=np.int32)
        ss = np.ascontiguousarray(s, dtype=np.float32)
        tt = np.ascontiguousarray(t, dtype=np.float32)
        out = np.zeros(len(pf), dtype=PATCH_COORD_DTYPE)
        self._L.ref_patch_map_find(pt.handle, len(pf), _p(pf), _p(ss), _p(tt), _p(out))
        return out


def far_basis(pt: PatchTable, coords: np.ndarray):
    n = len(coords)
    w = [np.zeros((n, 20), dtype=np.float32) for _ in range(6)]
    lib().ref_patch_table_far_basis(pt.handle, n, _p(coords), *[_p(x) for x in w])
    return w


def osd_patch_basis(patch_type: int, field0: int, field1: int, s: float, t: float, nw: int = 6):
    w = [np.zeros(20, dtype=np.float32) for _ in range(6)]
    ptrs = [_p(x) for x in w[:nw]] + [None] * (6 - nw)
    f0 = int(np.array(field0, dtype=np.uint32).view(np.int32))
    f1 = int(np.array(field1, dtype=np.uint32).view(np.int32))
    n = lib().ref_osd_patch_basis(patch_type, f0, f1, s, t, *ptrs)
    return n, w[:nw]


# -------------------------------------------------------------------------------- evaluators --
def _descs(descs: Sequence[Sequence[int]]) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(descs, dtype=np.int32).reshape(-1, 3))


def eval_stencils(src: np.ndarray, src_desc, dsts: Sequence[np.ndarray], dst_descs, table: StencilTable,
                  start: int = 0, end: Optional[int] = None, impl: str = "cpu") -> bool:
    """Osd::CpuEvaluator::EvalStencils (osd/cpuEvaluator.cpp:37-125) or the OpenMP twin; dsts are written in place."""
    nw = len(dsts)
    assert nw in (1, 3, 6)
    end = table.num_stencils if end is None else end
    sd = _descs([src_desc])
    dd = _descs(dst_descs)
    dptr = (C.c_void_p * nw)(*[d.ctypes.data for d in dsts])
    ws = table.weight_streams(nw)
    wptr = (C.c_void_p * nw)(*[w.ctypes.data for w in ws])
    r = lib().ref_eval_stencils({"cpu": 0, "omp": 1}[impl], nw, _p(src), _p(sd), dptr, _p(dd), _p(table.sizes),
                                _p(table.offsets), _p(table.indices), wptr, start, end)
    if r < 0:
        raise RuntimeError("reference evaluator unavailable: " + impl)
    return bool(r)


def eval_patches(src: np.ndarray, src_desc, dsts: Sequence[np.ndarray], dst_descs, coords: np.ndarray,
                 triple: PatchTriple, impl: str = "cpu") -> bool:
    """Osd::CpuEvaluator::EvalPatches (osd/cpuEvaluator.cpp:157-381) or the OpenMP twin."""
    nw = len(dsts)
    assert nw in (1, 3, 6)
    sd = _descs([src_desc])
    dd = _descs(dst_descs)
    dptr = (C.c_void_p * nw)(*[d.ctypes.data for d in dsts])
    r = lib().ref_eval_patches({"cpu": 0, "omp": 1}[impl], nw, _p(src), _p(sd), dptr, _p(dd), len(coords),
                               _p(coords), _p(triple.arrays), _p(triple.indices), _p(triple.params))
    if r < 0:
        raise RuntimeError("reference evaluator unavailable: " + impl)
    return bool(r)


def far_update_values_xyz(mesh: Mesh, src_xyz: np.ndarray, mode="vertex", intermediate_levels=False) -> np.ndarray:
    """Far::StencilTable::UpdateValues (far/stencilTable.h:648-674): third independent implementation."""
    L = lib()
    h = L.ref_stencil_table_create(mesh.h, 0, int(intermediate_levels), 1, 0, None)
    n = L.ref_stencil_table_num_stencils(h)
    src = np.ascontiguousarray(src_xyz, dtype=np.float32)
    dst = np.zeros((n, 3), dtype=np.float32)
    L.ref_stencil_table_update_values_xyz(h, _p(src), _p(dst))
    L.ref_stencil_table_free(h)
    return dst
