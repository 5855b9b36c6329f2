"""Synthetic deforming-mesh workloads in the reference's table formats (no reference code involved).

bench.py and the scale tests need multi-million-row stencil / patch tables on the GPU box, where the
reference (and therefore Far) is not available.  This module builds them from first principles for closed,
crease-free meshes:

  * uniform Catmull-Clark / Loop refinement as sparse subdivision matrices, composed level by level; the
    rows of the product ARE the factorised stencils Far::StencilTableFactory produces
    (far/stencilTableFactory.cpp:78-151), in Far's vertex order (vtr/refinement.cpp:240-272: children of
    vertices, then of faces, then of edges; child topology as vtr/quadRefinement.cpp / triRefinement.cpp);
  * for regular quad tori: the level-0 bicubic B-spline patch table (one REGULAR patch per face, the layout
    Osd::CpuPatchTable flattens, osd/cpuPatchTable.cpp:35-156) and limit-stencil tables (what
    Far::LimitStencilTableFactory yields on an all-regular mesh: 16 weights per location and derivative).

tests/test_synth_vs_far.py checks these generators against the real Far tables row by row.
Weights are computed in float64 and rounded once to float32 (Far accumulates in float32: <= 1e-7 apart).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import scipy.sparse as sp

from .osd import PATCH_ARRAY_DTYPE, PATCH_COORD_DTYPE, PATCH_PARAM_DTYPE


@dataclass
class SynthMesh:
    scheme: str                 # "catmark" (quads) or "loop" (triangles)
    positions: np.ndarray       # [V,3] float32
    faces: np.ndarray           # [F,4] or [F,3] int32
    grid: Optional[Tuple[int, int]] = None   # (nu, nv) for tori

    @property
    def num_verts(self) -> int:
        return int(self.positions.shape[0])


@dataclass
class SynthStencilTable:
    """Same attribute names as the Far accessors (GetSizes/GetOffsets/GetControlIndices/GetWeights...)."""
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

    @property
    def num_elements(self) -> int:
        return int(self.indices.shape[0])

    def weight_streams(self, nw: int):
        return [self.weights, self.du, self.dv, self.duu, self.duv, self.dvv][:nw]

    def algorithmic_bytes(self, n_out: int, length: int, src_stride: int) -> int:
        """sum_rows[8 + size*4*(1+K)] + nCV*srcStride*4 + nRows*K*L*4 (reference formats)."""
        return int(8 * self.num_stencils + 4 * (1 + n_out) * self.num_elements
                   + 4 * self.num_control_verts * src_stride + 4 * self.num_stencils * n_out * length)

    def row_range(self, start: int, end: int) -> "SynthStencilTable":
        """Rows [start,end) as a self-contained table (offsets re-based) -- used for row-range sharding."""
        e0 = int(self.offsets[start]) if start < self.num_stencils else self.num_elements
        e1 = int(self.offsets[end]) if end < self.num_stencils else self.num_elements
        cut = lambda a: None if a is None else a[e0:e1]
        return SynthStencilTable(self.num_control_verts, self.sizes[start:end], self.offsets[start:end] - e0,
                                 self.indices[e0:e1], self.weights[e0:e1], cut(self.du), cut(self.dv),
                                 cut(self.duu), cut(self.duv), cut(self.dvv))


# ------------------------------------------------------------------------------------- meshes --
def torus_positions(nu: int, nv: int, R: float = 3.0, r: float = 1.0) -> np.ndarray:
    u = (np.arange(nu, dtype=np.float64) * (2.0 * np.pi / nu))[:, None]
    v = (np.arange(nv, dtype=np.float64) * (2.0 * np.pi / nv))[None, :]
    x = (R + r * np.cos(v)) * np.cos(u)
    y = (R + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v) + 0.0 * u
    return np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)


def torus_quads(nu: int, nv: int) -> SynthMesh:
    """Closed quad torus: nu*nv vertices (all valence 4) and nu*nv faces; vertex (i,j) has index i*nv+j."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    i1, j1 = (i + 1) % nu, (j + 1) % nv
    faces = np.stack([i * nv + j, i1 * nv + j, i1 * nv + j1, i * nv + j1], axis=-1).reshape(-1, 4).astype(np.int32)
    return SynthMesh("catmark", torus_positions(nu, nv), faces, (nu, nv))


def torus_tris(nu: int, nv: int) -> SynthMesh:
    """Closed triangle torus: nu*nv vertices (all valence 6) and 2*nu*nv faces."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    i1, j1 = (i + 1) % nu, (j + 1) % nv
    a, b, c, d = i * nv + j, i1 * nv + j, i1 * nv + j1, i * nv + j1
    t0 = np.stack([a, b, c], axis=-1)
    t1 = np.stack([a, c, d], axis=-1)
    faces = np.stack([t0, t1], axis=2).reshape(-1, 3).astype(np.int32)
    return SynthMesh("loop", torus_positions(nu, nv), faces, (nu, nv))


def deform(positions: np.ndarray, frame: int) -> np.ndarray:
    """Per-frame deformation in the style of examples/glEvalLimit (rotate about z by an angle that depends on z)."""
    p = positions.astype(np.float32)
    ang = (p[:, 2] * np.float32(np.sin(0.1 * frame))).astype(np.float32)
    c, s = np.cos(ang), np.sin(ang)
    out = np.empty_like(p)
    out[:, 0] = p[:, 0] * c - p[:, 1] * s
    out[:, 1] = p[:, 0] * s + p[:, 1] * c
    out[:, 2] = p[:, 2]
    return out


def vertex_normals_like(positions: np.ndarray) -> np.ndarray:
    """A cheap unit 'normal' primvar (direction from the torus centre line); only used as 3 more floats/vertex."""
    p = positions.astype(np.float64)
    ring = p.copy()
    ring[:, 2] = 0.0
    nrm = np.linalg.norm(ring, axis=1, keepdims=True)
    ring = ring / np.maximum(nrm, 1e-12) * 3.0
    n = p - ring
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
    return n.astype(np.float32)


# ------------------------------------------------------------------------------- topology ----
def _first_encounter_edges(faces: np.ndarray, nV: int):
    """Edges numbered in first-encounter order over faces / face-edges (vtr/level.cpp:1640-1750)."""
    F, n = faces.shape
    v0 = faces.reshape(-1).astype(np.int64)
    v1 = np.roll(faces, -1, axis=1).reshape(-1).astype(np.int64)
    key = np.minimum(v0, v1) * nV + np.maximum(v0, v1)
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")           # unique keys sorted by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    face_edges = rank[inv].reshape(F, n).astype(np.int32)
    fo = first[order]
    edges = np.stack([v0[fo], v1[fo]], axis=1).astype(np.int32)
    return edges, face_edges


def _lookup_face_edges(faces: np.ndarray, edges: np.ndarray, nV: int) -> np.ndarray:
    ek = np.minimum(edges[:, 0], edges[:, 1]).astype(np.int64) * nV + np.maximum(edges[:, 0], edges[:, 1])
    order = np.argsort(ek, kind="stable")
    v0 = faces.astype(np.int64)
    v1 = np.roll(faces, -1, axis=1).astype(np.int64)
    fk = np.minimum(v0, v1) * nV + np.maximum(v0, v1)
    pos = np.searchsorted(ek[order], fk.reshape(-1))
    return order[pos].reshape(faces.shape).astype(np.int32)


def _refine_catmark(faces, edges, face_edges, nV):
    F, E = faces.shape[0], edges.shape[0]
    vv = np.arange(nV, dtype=np.int64)
    fv = nV + np.arange(F, dtype=np.int64)              # children of vertices, faces, edges (default Far order)
    ev = nV + F + np.arange(E, dtype=np.int64)
    new_faces = np.empty((F, 4, 4), dtype=np.int64)
    for j in range(4):
        jn, jo, jp = (j + 1) % 4, (j + 2) % 4, (j + 3) % 4
        new_faces[:, j, j] = vv[faces[:, j]]
        new_faces[:, j, jn] = ev[face_edges[:, j]]
        new_faces[:, j, jo] = fv
        new_faces[:, j, jp] = ev[face_edges[:, jp]]
    new_faces = new_faces.reshape(4 * F, 4)
    e_face = np.stack([np.repeat(fv, 4), ev[face_edges.reshape(-1)]], axis=1)               # 4f+j
    e_edge = np.stack([np.repeat(ev, 2), vv[edges.reshape(-1)]], axis=1)                     # 4F+2e+j
    new_edges = np.concatenate([e_face, e_edge], axis=0)
    nV2 = nV + F + E

    # subdivision matrix rows: [vertex points | face points | edge points]
    valence = np.bincount(edges.reshape(-1), minlength=nV).astype(np.float64)
    rows, cols, vals = [], [], []
    # face points
    rows.append(np.repeat(fv, 4)); cols.append(faces.reshape(-1)); vals.append(np.full(4 * F, 0.25))
    # edge points: (v0 + v1)/4 + (sum of the verts of both adjacent faces)/16
    rows.append(np.repeat(ev, 2)); cols.append(edges.reshape(-1)); vals.append(np.full(2 * E, 0.25))
    he_edge = face_edges.reshape(-1)                                                         # half-edge -> edge
    rows.append(np.repeat(ev[he_edge], 4)); cols.append(np.repeat(faces, 4, axis=0).reshape(-1))
    vals.append(np.full(16 * F, 1.0 / 16.0))
    # vertex points: (n-2)/n v + 1/n^2 sum(edge neighbours) + 1/n^2 sum(face points)
    rows.append(vv); cols.append(vv); vals.append((valence - 2.0) / valence)
    a, b = edges[:, 0].astype(np.int64), edges[:, 1].astype(np.int64)
    rows.append(a); cols.append(b); vals.append(1.0 / valence[a] ** 2)
    rows.append(b); cols.append(a); vals.append(1.0 / valence[b] ** 2)
    fvert = faces.reshape(-1).astype(np.int64)
    rows.append(np.repeat(fvert, 4)); cols.append(np.repeat(faces, 4, axis=0).reshape(-1))
    vals.append(np.repeat(0.25 / valence[fvert] ** 2, 4))
    S = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nV2, nV)).tocsr()
    new_face_edges = _lookup_face_edges(new_faces, new_edges, nV2)
    return new_faces.astype(np.int32), new_edges.astype(np.int32), new_face_edges, nV2, S


def _refine_loop(faces, edges, face_edges, nV):
    F, E = faces.shape[0], edges.shape[0]
    vv = np.arange(nV, dtype=np.int64)
    ev = nV + np.arange(E, dtype=np.int64)
    e0, e1, e2 = ev[face_edges[:, 0]], ev[face_edges[:, 1]], ev[face_edges[:, 2]]
    f0, f1, f2 = vv[faces[:, 0]], vv[faces[:, 1]], vv[faces[:, 2]]
    new_faces = np.stack([np.stack([f0, e0, e2], 1), np.stack([e0, f1, e1], 1),
                          np.stack([e2, e1, f2], 1), np.stack([e1, e2, e0], 1)], axis=1).reshape(4 * F, 3)
    e_face = np.stack([np.stack([e0, e2], 1), np.stack([e1, e0], 1), np.stack([e2, e1], 1)], axis=1).reshape(3 * F, 2)
    e_edge = np.stack([np.repeat(ev, 2), vv[edges.reshape(-1)]], axis=1)
    new_edges = np.concatenate([e_face, e_edge], axis=0)
    nV2 = nV + E

    valence = np.bincount(edges.reshape(-1), minlength=nV).astype(np.float64)
    beta = 0.25 * np.cos(2.0 * np.pi / valence) + 0.375                     # sdc/loopScheme.h:180-215
    ew = (0.625 - beta * beta) / valence
    ew = np.where(valence == 6, 0.0625, ew)
    vw = np.where(valence == 6, 0.625, 1.0 - ew * valence)
    rows, cols, vals = [], [], []
    rows.append(vv); cols.append(vv); vals.append(vw)
    a, b = edges[:, 0].astype(np.int64), edges[:, 1].astype(np.int64)
    rows.append(a); cols.append(b); vals.append(ew[a])
    rows.append(b); cols.append(a); vals.append(ew[b])
    # edge points: 3/8 (v0+v1) + 1/8 (vertex opposite the edge in each adjacent face)   sdc/loopScheme.h:84-125
    rows.append(np.repeat(ev, 2)); cols.append(edges.reshape(-1)); vals.append(np.full(2 * E, 0.375))
    opp = np.roll(faces, -2, axis=1)                                        # vertex opposite face-edge i is v[i+2]
    rows.append(ev[face_edges.reshape(-1)]); cols.append(opp.reshape(-1)); vals.append(np.full(3 * F, 0.125))
    S = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nV2, nV)).tocsr()
    new_face_edges = _lookup_face_edges(new_faces, new_edges, nV2)
    return new_faces.astype(np.int32), new_edges.astype(np.int32), new_face_edges, nV2, S


def uniform_stencil_table(mesh: SynthMesh, level: int, return_topology: bool = False):
    """Factorised stencils of the LAST level of `level` uniform refinements (generateIntermediateLevels=false,
    the Osd::Mesh default for uniform refinement, osd/mesh.h:588-591)."""
    assert level >= 1
    nV = mesh.num_verts
    faces = mesh.faces.astype(np.int32)
    edges, face_edges = _first_encounter_edges(faces, nV)
    step = _refine_catmark if mesh.scheme == "catmark" else _refine_loop
    total = None
    n = nV
    for _ in range(level):
        faces, edges, face_edges, n, S = step(faces, edges, face_edges, n)
        total = S if total is None else (S @ total)
    total = total.tocsr()
    total.sort_indices()
    sizes = np.diff(total.indptr).astype(np.int32)
    table = SynthStencilTable(num_control_verts=nV, sizes=sizes, offsets=total.indptr[:-1].astype(np.int32),
                              indices=total.indices.astype(np.int32), weights=total.data.astype(np.float32))
    if return_topology:
        return table, faces
    return table


# --------------------------------------------------------------- regular-torus patch / limit tables --
def _bspline_1d(t: np.ndarray):
    t = t.astype(np.float64)
    t2, t3 = t * t, t * t * t
    b = np.stack([(1 - t) ** 3, 3 * t3 - 6 * t2 + 4, -3 * t3 + 3 * t2 + 3 * t + 1, t3], axis=-1) / 6.0
    d = np.stack([-(1 - t) ** 2, 3 * t2 - 4 * t, -3 * t2 + 2 * t + 1, t2], axis=-1) / 2.0
    dd = np.stack([1 - t, 3 * t - 2, -3 * t + 1, t], axis=-1)
    return b, d, dd


def torus_patch_cvs(nu: int, nv: int) -> np.ndarray:
    """[F,16] control vertices of the level-0 regular patch of every face: point 4*row+col, col along s (i), row along t (j)."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    cvs = np.empty((nu, nv, 16), dtype=np.int32)
    for row in range(4):
        for col in range(4):
            cvs[:, :, 4 * row + col] = ((i - 1 + col) % nu) * nv + ((j - 1 + row) % nv)
    return cvs.reshape(-1, 16)


@dataclass
class SynthPatchTriple:
    arrays: np.ndarray
    indices: np.ndarray
    params: np.ndarray


@dataclass
class SynthPatchTable:
    vertex: SynthPatchTriple
    varying: Optional[SynthPatchTriple] = None
    fvar: Optional[list] = None


def torus_patch_table(mesh: SynthMesh) -> SynthPatchTable:
    """One REGULAR (type 6) depth-0 patch per face + the QUADS (type 3) varying patches (far/patchTable.cpp:431-467)."""
    nu, nv = mesh.grid
    F = nu * nv
    arrays = np.zeros(1, dtype=PATCH_ARRAY_DTYPE)
    arrays[0] = (6, 6, F, 0, 16, 0)
    params = np.zeros(F, dtype=PATCH_PARAM_DTYPE)
    params["field0"] = np.arange(F, dtype=np.uint32)               # faceId, transition 0
    params["field1"] = np.uint32(1 << 5)                           # depth 0, regular, no boundary, u=v=0
    vertex = SynthPatchTriple(arrays, torus_patch_cvs(nu, nv).reshape(-1), params)
    varr = np.zeros(1, dtype=PATCH_ARRAY_DTYPE)
    varr[0] = (3, 3, F, 0, 4, 0)
    varying = SynthPatchTriple(varr, mesh.faces.reshape(-1).astype(np.int32), params)
    return SynthPatchTable(vertex=vertex, varying=varying, fvar=[])


def random_patch_coords(num_patches: int, n: int, seed: int = 2024, sort_by_patch: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    coords = np.zeros(n, dtype=PATCH_COORD_DTYPE)
    p = rng.integers(0, num_patches, size=n, dtype=np.int64)
    if sort_by_patch:
        p.sort()
    coords["arrayIndex"] = 0
    coords["patchIndex"] = p
    coords["vertIndex"] = p * 16
    coords["s"] = rng.random(n, dtype=np.float32)
    coords["t"] = rng.random(n, dtype=np.float32)
    return coords


def torus_limit_stencil_table(mesh: SynthMesh, face: np.ndarray, s: np.ndarray, t: np.ndarray,
                              second: bool = True) -> SynthStencilTable:
    """Limit stencils with 1st (and 2nd) derivative weights at (face, s, t) on the all-regular torus: the 16 bicubic
    B-spline weights of the face's patch -- what Far::LimitStencilTableFactory::Create computes there
    (far/stencilTableFactory.cpp:559-635)."""
    nu, nv = mesh.grid
    cvs = torus_patch_cvs(nu, nv)[np.asarray(face, dtype=np.int64)]
    bs, ds, dss = _bspline_1d(np.asarray(s))
    bt, dt, dtt = _bspline_1d(np.asarray(t))
    outer = lambda ct, cs: (ct[:, :, None] * cs[:, None, :]).reshape(len(ct), 16).astype(np.float32).reshape(-1)
    n = len(cvs)
    tbl = SynthStencilTable(num_control_verts=mesh.num_verts, sizes=np.full(n, 16, dtype=np.int32),
                            offsets=(np.arange(n, dtype=np.int64) * 16).astype(np.int32),
                            indices=cvs.reshape(-1).astype(np.int32), weights=outer(bt, bs),
                            du=outer(bt, ds), dv=outer(dt, bs))
    if second:
        tbl.duu, tbl.duv, tbl.dvv = outer(bt, dss), outer(dt, ds), outer(dtt, bs)
    return tbl
