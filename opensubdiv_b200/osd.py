"""Host-side mirror of the Osd evaluator interface for the B200 backend.

Same class / method names, argument meaning and error behaviour as the reference CUDA backend
(paths relative to /root/reference/opensubdiv):

    BufferDescriptor   osd/bufferDescriptor.h:61-104
    B200VertexBuffer   <-> CudaVertexBuffer   osd/cudaVertexBuffer.h:42-80
    B200StencilTable   <-> CudaStencilTable   osd/cudaEvaluator.h:52-92
    B200PatchTable     <-> CudaPatchTable     osd/cudaPatchTable.h:51-112
    B200Evaluator      <-> CudaEvaluator      osd/cudaEvaluator.h:94-1262

Everything here is a thin veneer over the C ABI (include/b200osd_capi.h); the C++ header-only twin that
drops into Osd::Mesh<> lives in include/b200osd/.  PyTorch is used for plumbing only (current stream,
zero-copy tensor views for torch.distributed).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import capi

# numpy mirrors of the Osd POD types
PATCH_COORD_DTYPE = np.dtype([("arrayIndex", "<i4"), ("patchIndex", "<i4"), ("vertIndex", "<i4"),
                              ("s", "<f4"), ("t", "<f4")])                               # osd/types.h:42-64
PATCH_ARRAY_DTYPE = np.dtype([("regDesc", "<i4"), ("desc", "<i4"), ("numPatches", "<i4"),
                              ("indexBase", "<i4"), ("stride", "<i4"), ("primitiveIdBase", "<i4")])  # :66-122
PATCH_PARAM_DTYPE = np.dtype([("field0", "<u4"), ("field1", "<u4"), ("sharpness", "<f4")])           # :127-130
PATCH_GREGORY_TRUE_DERIVATIVES = 1          # B200OSD_PATCH_GREGORY_TRUE_DERIVATIVES (include/b200osd_capi.h)


@dataclass(frozen=True)
class BufferDescriptor:
    """offset / length / stride in floats (osd/bufferDescriptor.h:61-104)."""
    offset: int = 0
    length: int = 0
    stride: int = 0

    def GetLocalOffset(self) -> int:
        return self.offset % self.stride if self.stride > 0 else 0

    def IsValid(self) -> bool:
        return self.length > 0 and self.length <= self.stride - self.GetLocalOffset()

    def as_c(self):
        return (C.c_int * 3)(self.offset, self.length, self.stride)


def _desc(d) -> BufferDescriptor:
    return d if isinstance(d, BufferDescriptor) else BufferDescriptor(*d)


def _stream_ptr(deviceContext=None) -> Optional[int]:
    """deviceContext may be None (torch's current stream), an int cudaStream_t, or a torch.cuda.Stream."""
    if deviceContext is None:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.cuda.current_stream().cuda_stream or None
        except Exception:
            pass
        return None
    if isinstance(deviceContext, int):
        return deviceContext or None
    return getattr(deviceContext, "cuda_stream", None) or None


def _dev_ptr(buf) -> Optional[int]:
    if buf is None:
        return None
    if hasattr(buf, "BindCudaBuffer"):
        return buf.BindCudaBuffer()
    if hasattr(buf, "data_ptr"):
        if not buf.is_cuda:
            raise capi.B200OsdError("evaluator buffers must live on the GPU (got a CPU tensor)")
        return buf.data_ptr()
    if isinstance(buf, int):
        return buf
    raise TypeError(f"cannot take a device pointer from {type(buf)}")


class _CudaArrayView:
    def __init__(self, ptr: int, n: int, owner, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        self._owner = owner


class B200VertexBuffer:
    """Device vertex buffer; mirrors CudaVertexBuffer (Create / UpdateData / GetNum* / BindCudaBuffer)."""

    def __init__(self, handle, numElements: int, numVertices: int):
        self._h = handle
        self._numElements = numElements
        self._numVertices = numVertices
        self._tensor = None

    @classmethod
    def Create(cls, numElements: int, numVertices: int, deviceContext=None) -> Optional["B200VertexBuffer"]:
        h = capi.lib().b200osd_vertex_buffer_create(numElements, numVertices)
        return cls(h, numElements, numVertices) if h else None       # NULL on failure, like the reference

    def __del__(self):
        try:
            if self._h:
                capi.lib().b200osd_vertex_buffer_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def UpdateData(self, src, startVertex: int, numVertices: int, deviceContext=None) -> None:
        """Host -> device copy of numVertices vertices (src: numpy array or CPU tensor, ideally pinned)."""
        ptr = src.ctypes.data if isinstance(src, np.ndarray) else src.data_ptr()
        rc = capi.lib().b200osd_vertex_buffer_update(self._h, ptr, startVertex, numVertices, _stream_ptr(deviceContext))
        if not capi.check(rc, "B200VertexBuffer::UpdateData"):
            raise capi.B200OsdError(capi.last_error())

    def ReadData(self, dst, startVertex: int, numVertices: int, deviceContext=None) -> None:
        """Device -> host read-back into dst (numpy array or CPU tensor); asynchronous on the stream."""
        ptr = dst.ctypes.data if isinstance(dst, np.ndarray) else dst.data_ptr()
        rc = capi.lib().b200osd_vertex_buffer_read(self._h, ptr, startVertex, numVertices, _stream_ptr(deviceContext))
        if not capi.check(rc, "B200VertexBuffer::ReadData"):
            raise capi.B200OsdError(capi.last_error())

    def GetNumElements(self) -> int:
        return self._numElements

    def GetNumVertices(self) -> int:
        return self._numVertices

    def BindCudaBuffer(self) -> int:
        return capi.lib().b200osd_vertex_buffer_bind(self._h)

    def BindVBO(self, deviceContext=None) -> int:      # Osd::Mesh calls this name (osd/mesh.h:562-568)
        return self.BindCudaBuffer()

    def as_tensor(self):
        """Zero-copy torch view [numVertices, numElements] (for torch.distributed collectives)."""
        if self._tensor is None:
            import torch
            view = _CudaArrayView(self.BindCudaBuffer(), self._numElements * self._numVertices, self)
            self._tensor = torch.as_tensor(view, device="cuda").view(self._numVertices, self._numElements)
        return self._tensor


class B200StencilTable:
    """Device stencil table; mirrors CudaStencilTable plus the B200 bucketed layout."""

    def __init__(self, handle, keep=None):
        self._h = handle
        self._keep = keep

    @classmethod
    def Create(cls, table, deviceContext=None, bucketed: bool = True, locality: bool = False,
               idx16: bool = True, sort_elements: bool = False, keep_order: bool = False,
               host_layout: bool = False, reference_exact: bool = False) -> Optional["B200StencilTable"]:
        """`table` is anything with the Far::StencilTable / LimitStencilTable accessors as numpy arrays:
        sizes, offsets, indices, weights and optionally du, dv, duu, duv, dvv (far/stencilTable.h:156-186,434-456).
        Summation order (include/b200osd_capi.h): default = rows of <= 16 terms in control-index order; keep_order = the
        table's own order everywhere; sort_elements = every row in control-index order.  host_layout = build the bucketed
        copy on the host instead of on the device (the same bytes).  reference_exact = evaluate with the CPU reference's own
        arithmetic (rounded product, then add, table order): bit-identical to Osd::CpuEvaluator, about half the speed."""
        def arr(name, dt):
            a = getattr(table, name, None)
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        sizes, offsets, indices = arr("sizes", np.int32), arr("offsets", np.int32), arr("indices", np.int32)
        w = [arr(n, np.float32) for n in ("weights", "du", "dv", "duu", "duv", "dvv")]
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        ncv = int(getattr(table, "num_control_verts", 0) or 0)       # Far::StencilTable::GetNumControlVertices(); 0 = derive
        h = capi.lib().b200osd_stencil_table_create(len(sizes), ncv, p(sizes), p(offsets), p(indices), *[p(x) for x in w],
                                                    (0 if bucketed else 1) | (2 if locality else 0) | (0 if idx16 else 4)
                                                    | (8 if sort_elements else 0) | (16 if keep_order else 0) | (32 if host_layout else 0) | (64 if reference_exact else 0))
        return cls(h) if h else None

    @classmethod
    def CreateFromDevice(cls, numStencils: int, sizes, offsets, indices, weights, du=None, dv=None, duu=None, duv=None,
                         dvv=None, numControlVertices: int = 0, keep_order: bool = False) -> Optional["B200StencilTable"]:
        """From DEVICE arrays in the reference layout (the buffers of an Osd::CudaStencilTable, osd/cudaEvaluator.h:57-90):
        one conversion to the bucketed layout instead of evaluating the raw arrays row by row (EvalStencilsRaw)."""
        p = lambda a: None if a is None else _dev_ptr(a)
        h = capi.lib().b200osd_stencil_table_create_from_device(int(numStencils), int(numControlVertices), p(sizes), p(offsets),
                                                                p(indices), p(weights), p(du), p(dv), p(duu), p(duv), p(dvv),
                                                                16 if keep_order else 0)
        if not h:
            raise capi.B200OsdError("B200StencilTable::CreateFromDevice: " + capi.last_error())
        return cls(h)

    @classmethod
    def CreateLimitStencils(cls, patchTable, cvStencilTable, numLocations: int, patchCoords, numWeightSets: int = 6,
                            bucketed: bool = True, deviceContext=None) -> Optional["B200StencilTable"]:
        """Far::LimitStencilTableFactory::Create on the DEVICE (far/stencilTableFactory.cpp:559-662): patchCoords are located
        samples (device PatchCoord records, e.g. from B200PatchMap.FindPatches), cvStencilTable the refined + local-point
        stencil table of the topology; numWeightSets 1 / 3 / 6 = value / + 1st / + 2nd derivative weights."""
        h = capi.lib().b200osd_limit_stencil_table_create(patchTable._h, cvStencilTable._h, int(numLocations), _dev_ptr(patchCoords),
                                                          int(numWeightSets), 0 if bucketed else 1, _stream_ptr(deviceContext))
        if not h:
            raise capi.B200OsdError("B200StencilTable::CreateLimitStencils: " + capi.last_error())
        return cls(h)

    def __del__(self):
        try:
            if self._h:
                capi.lib().b200osd_stencil_table_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _buf(self, which):
        return capi.lib().b200osd_stencil_table_buffer(self._h, which)

    def GetSizesBuffer(self): return self._buf(0)
    def GetOffsetsBuffer(self): return self._buf(1)
    def GetIndicesBuffer(self): return self._buf(2)
    def GetWeightsBuffer(self): return self._buf(3)
    def GetDuWeightsBuffer(self): return self._buf(4)
    def GetDvWeightsBuffer(self): return self._buf(5)
    def GetDuuWeightsBuffer(self): return self._buf(6)
    def GetDuvWeightsBuffer(self): return self._buf(7)
    def GetDvvWeightsBuffer(self): return self._buf(8)

    def GetNumStencils(self) -> int:
        return capi.lib().b200osd_stencil_table_num_stencils(self._h)

    def GetNumControlVertices(self) -> int:
        return capi.lib().b200osd_stencil_table_num_control_vertices(self._h)

    def GetNumElements(self) -> int:
        return capi.lib().b200osd_stencil_table_num_elements(self._h)

    def GetStreamBytes(self, nOut: int = 1) -> int:
        return capi.lib().b200osd_stencil_table_stream_bytes(self._h, nOut)

    def ToHost(self, numWeightSets: int = 1):
        """(sizes, offsets, indices, [weights, du, ...]) of the reference-layout device arrays as numpy arrays."""
        import torch
        n, ne = self.GetNumStencils(), self.GetNumElements()

        def get(ptr, count, typestr):
            if not ptr or count == 0:
                return np.zeros(0, np.int32 if typestr == "<i4" else np.float32)
            return torch.as_tensor(_CudaArrayView(ptr, count, self, typestr), device="cuda").cpu().numpy().copy()
        return (get(self._buf(0), n, "<i4"), get(self._buf(1), n, "<i4"), get(self._buf(2), ne, "<i4"),
                [get(self._buf(3 + k), ne, "<f4") for k in range(numWeightSets)])

    def IsFactorized(self) -> bool:
        """False for a table built with factorizeIntermediateLevels = false (apply it one level at a time)."""
        return bool(capi.lib().b200osd_stencil_table_is_factorized(self._h))

    def SetVariant(self, variant: int) -> None:
        """Kernel variant for this table (bench / tests): 0 auto, see include/b200osd_capi.h."""
        capi.lib().b200osd_stencil_table_set_variant(self._h, int(variant))

    def GetVariant(self) -> int:
        return capi.lib().b200osd_stencil_table_get_variant(self._h)


class B200PatchTable:
    """Device patch table; mirrors CudaPatchTable (built from the flattened Osd::CpuPatchTable arrays)."""

    VertexBufferBinding = int     # Osd::Mesh needs PT::VertexBufferBinding (osd/mesh.h:71,426)

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def Create(cls, table, deviceContext=None) -> Optional["B200PatchTable"]:
        """`table` has .vertex, .varying (or None) and .fvar (list) triples, each with numpy .arrays
        (PATCH_ARRAY_DTYPE), .indices (int32) and .params (PATCH_PARAM_DTYPE) -- the layout Osd::CpuPatchTable
        produces from a Far::PatchTable (osd/cpuPatchTable.cpp:35-156)."""
        L = capi.lib()
        fvar = list(getattr(table, "fvar", []) or [])
        h = L.b200osd_patch_table_create(len(fvar))
        if not h:
            return None
        self = cls(h)

        def put(which, tr, with_params=True):
            if tr is None:
                return True
            a = np.ascontiguousarray(tr.arrays, dtype=PATCH_ARRAY_DTYPE)
            ix = np.ascontiguousarray(tr.indices, dtype=np.int32)
            pr = np.ascontiguousarray(tr.params, dtype=PATCH_PARAM_DTYPE) if with_params else None
            rc = L.b200osd_patch_table_set(h, which, len(a), a.ctypes.data, len(ix), ix.ctypes.data if len(ix) else None,
                                           0 if pr is None else len(pr), None if pr is None else pr.ctypes.data)
            return capi.check(rc, "B200PatchTable::Create")
        ok = put(0, table.vertex) and put(1, getattr(table, "varying", None), with_params=False)
        for c, tr in enumerate(fvar):
            ok = ok and put(2 + c, tr)
        return self if ok else None

    def __del__(self):
        try:
            if self._h:
                capi.lib().b200osd_patch_table_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _buf(self, which, kind):
        return capi.lib().b200osd_patch_table_buffer(self._h, which, kind)

    def GetPatchArrayBuffer(self): return self._buf(0, 0)
    def GetPatchIndexBuffer(self): return self._buf(0, 1)
    def GetPatchParamBuffer(self): return self._buf(0, 2)
    def GetVaryingPatchArrayBuffer(self): return self._buf(1, 0)
    def GetVaryingPatchIndexBuffer(self): return self._buf(1, 1)
    def GetNumFVarChannels(self) -> int: return capi.lib().b200osd_patch_table_num_fvar_channels(self._h)
    def GetFVarPatchArrayBuffer(self, fvarChannel: int = 0): return self._buf(2 + fvarChannel, 0)
    def GetFVarPatchIndexBuffer(self, fvarChannel: int = 0): return self._buf(2 + fvarChannel, 1)
    def GetFVarPatchParamBuffer(self, fvarChannel: int = 0): return self._buf(2 + fvarChannel, 2)

    def SetVariant(self, variant: int) -> None:
        """How EvalPatches* calls through this table are served: 0 automatic, 1 caller's order, 2 grouped by patch per call,
        3 per-call hull cache (include/b200osd_capi.h)."""
        capi.lib().b200osd_patch_table_set_variant(self._h, int(variant))

    def GetVariant(self) -> int:
        return capi.lib().b200osd_patch_table_get_variant(self._h)

    def SetGregoryTrueDerivatives(self, on: bool = True) -> None:
        """The reference's build option OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES (osd/patchBasis.h:421-487) as a per-table
        run-time option: derivative weights of the interior points of GREGORY_BASIS patches by quotient + product rule
        instead of the default approximation.  Applies to EvalPatches* through this table and to CreateLimitStencils."""
        if not capi.check(capi.lib().b200osd_patch_table_set_options(self._h, PATCH_GREGORY_TRUE_DERIVATIVES if on else 0),
                          "B200PatchTable::SetGregoryTrueDerivatives"):
            raise capi.B200OsdError(capi.last_error())

    def GetGregoryTrueDerivatives(self) -> bool:
        return bool(capi.lib().b200osd_patch_table_get_options(self._h) & PATCH_GREGORY_TRUE_DERIVATIVES)


class B200PatchMap:
    """Device-resident Far::PatchMap (far/patchMap.h:48-217): locates (ptexFace, s, t) samples in the patches of a
    table and emits Osd::PatchCoord records on the device, ready for EvalPatches.  The reference answers one sample
    per FindPatch call on the host (examples/glEvalLimit/particles.cpp:91-115,392-394); here a whole batch is one
    kernel launch and never leaves HBM."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def Create(cls, table, patchesAreTriangular: Optional[bool] = None) -> Optional["B200PatchMap"]:
        """`table`: the same flattened patch-table object B200PatchTable.Create takes (.vertex.arrays/.params, and
        .varying for the triangular test of far/patchMap.cpp:93-94 unless patchesAreTriangular is given)."""
        a = np.ascontiguousarray(table.vertex.arrays, dtype=PATCH_ARRAY_DTYPE)
        pr = np.ascontiguousarray(table.vertex.params, dtype=PATCH_PARAM_DTYPE)
        if patchesAreTriangular is None:
            var = getattr(table, "varying", None)
            patchesAreTriangular = bool(var is not None and len(var.arrays) and int(var.arrays["desc"][0]) == 4)
        h = capi.lib().b200osd_patch_map_create(len(a), a.ctypes.data if len(a) else None, len(pr),
                                                pr.ctypes.data if len(pr) else None, int(patchesAreTriangular))
        if not h:
            raise capi.B200OsdError("B200PatchMap::Create failed: " + capi.last_error())
        return cls(h)

    def __del__(self):
        try:
            if self._h:
                capi.lib().b200osd_patch_map_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _info(self):
        info = (C.c_int * 6)()
        capi.check(capi.lib().b200osd_patch_map_info(self._h, info), "B200PatchMap::info")
        return list(info)

    def GetMinPatchFace(self) -> int: return self._info()[0]
    def GetMaxPatchFace(self) -> int: return self._info()[1]
    def GetMaxDepth(self) -> int: return self._info()[2]
    def GetNumNodes(self) -> int: return self._info()[4]
    def GetNumPatches(self) -> int: return self._info()[5]

    def FindPatches(self, numSamples: int, ptexFace, s, t, patchCoords, numFound=None, strides=(1, 1, 1),
                    deviceContext=None) -> bool:
        """ptexFace (int32), s, t (float32): DEVICE arrays with element strides `strides`; patchCoords: DEVICE buffer of
        numSamples 20-byte Osd::PatchCoord records.  A sample outside every patch (FindPatch == NULL) gets
        handle.arrayIndex = -1; B200Evaluator.EvalPatches* leaves its outputs untouched.  numFound: optional device int."""
        rc = capi.lib().b200osd_patch_map_find(self._h, numSamples, _dev_ptr(ptexFace), strides[0], _dev_ptr(s), strides[1],
                                               _dev_ptr(t), strides[2], _dev_ptr(patchCoords), _dev_ptr(numFound),
                                               _stream_ptr(deviceContext))
        return capi.check(rc, "B200PatchMap::FindPatches")


class B200FrameGraph:
    """A whole evaluation frame recorded once and replayed as ONE launch (no reference counterpart).

        frame = B200FrameGraph.Create()
        run_frame(deviceContext=frame)          # once eagerly on the frame's stream: first calls allocate
        frame.Synchronize()
        frame.Begin(); run_frame(deviceContext=frame); frame.End()
        ... update the control points in place ...; frame.Launch(); frame.Synchronize()

    Pass the object itself as `deviceContext` to the B200 classes (it exposes `.cuda_stream`)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def Create(cls) -> Optional["B200FrameGraph"]:
        h = capi.lib().b200osd_frame_create()
        return cls(h) if h else None

    def __del__(self):
        try:
            if self._h:
                capi.lib().b200osd_frame_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def cuda_stream(self) -> int:
        return capi.lib().b200osd_frame_stream(self._h) or 0

    @property
    def side_stream(self) -> int:
        """The frame's second (high-priority) stream: work issued on it while recording is a parallel branch."""
        return capi.lib().b200osd_frame_side_stream(self._h) or 0

    def Fence(self, mainWaitsForSide: bool) -> bool:
        return capi.check(capi.lib().b200osd_frame_fence(self._h, int(bool(mainWaitsForSide))), "B200FrameGraph::Fence")

    def SetL2Window(self, buf, num_bytes: int, hitRatio: float = 1.0) -> bool:
        """Keep [buf, buf + num_bytes) L2-resident across the frame's kernels (call before Begin); buf=None clears."""
        ptr = None if buf is None else (buf if isinstance(buf, int) else _dev_ptr(buf))
        return capi.check(capi.lib().b200osd_frame_set_l2_window(self._h, ptr, int(num_bytes), float(hitRatio)),
                          "B200FrameGraph::SetL2Window")

    def Begin(self) -> bool: return capi.check(capi.lib().b200osd_frame_begin(self._h), "B200FrameGraph::Begin")
    def End(self) -> bool: return capi.check(capi.lib().b200osd_frame_end(self._h), "B200FrameGraph::End")
    def Launch(self) -> bool: return capi.check(capi.lib().b200osd_frame_launch(self._h), "B200FrameGraph::Launch")
    def Synchronize(self) -> bool: return capi.check(capi.lib().b200osd_frame_synchronize(self._h), "B200FrameGraph::Synchronize")


def _is_desc(d) -> bool:
    return isinstance(d, BufferDescriptor) or (isinstance(d, (tuple, list)) and len(d) == 3
                                               and all(isinstance(v, (int, np.integer)) for v in d))


def _split_outputs(args):
    """(buf, desc, buf, desc, ..., rest...) -> ([(buf, desc)...], rest)."""
    outs, i = [], 0
    while i + 1 < len(args) and _is_desc(args[i + 1]):
        outs.append((args[i], _desc(args[i + 1])))
        i += 2
    return outs, args[i:]


class B200EvaluatorInstance:
    """What B200Evaluator.Create returns: the "instantiatable" flavour of an Osd evaluator (osd/mesh.h:305-409,
    osd/glComputeEvaluator.h:98-128).  Nothing needs compiling per descriptor set; the state worth caching is the
    grouping of one PatchCoord set by patch.  Pass the object as `instance` to the static EvalPatches* methods."""

    def __init__(self, descs):
        self.descs = descs
        self._plan = None
        self._table = None
        self._coords_ptr = None
        self._count = 0

    def __del__(self):
        try:
            if self._plan:
                capi.lib().b200osd_patch_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def BindPatchCoords(self, numPatchCoords: int, patchCoords, patchTable, deviceContext=None) -> bool:
        """Groups the coordinates by patch on the device and keeps the grouping for later EvalPatches* calls on the same
        buffer and table (any slot: vertex, varying, face-varying).  Call again after the coordinates change."""
        L = capi.lib()
        if self._plan is None or self._table is not patchTable or L.b200osd_patch_plan_capacity(self._plan) < numPatchCoords:
            if self._plan:
                L.b200osd_patch_plan_destroy(self._plan)
            self._plan = L.b200osd_patch_plan_create(patchTable._h, numPatchCoords)
            self._table = patchTable
            if not self._plan:
                raise capi.B200OsdError("B200Evaluator::BindPatchCoords: " + capi.last_error())
        self._coords_ptr = _dev_ptr(patchCoords)
        self._count = numPatchCoords
        return capi.check(L.b200osd_patch_plan_bin(self._plan, numPatchCoords, self._coords_ptr, _stream_ptr(deviceContext)),
                          "B200Evaluator::BindPatchCoords")

    def _matches(self, patchTable, numPatchCoords, patchCoords) -> bool:
        return (self._plan is not None and self._table is patchTable and self._count == numPatchCoords
                and self._coords_ptr == _dev_ptr(patchCoords))


class B200Evaluator:
    """Static evaluator; mirrors CudaEvaluator.  Every method returns the reference's bool."""

    Instantiatable = True

    @staticmethod
    def Create(srcDesc, dstDesc, duDesc=None, dvDesc=None, duuDesc=None, duvDesc=None, dvvDesc=None,
               deviceContext=None) -> B200EvaluatorInstance:
        """EVALUATOR::Create(srcDesc, dstDesc, duDesc, dvDesc[, duuDesc, duvDesc, dvvDesc], deviceContext) (osd/mesh.h:268-283)."""
        return B200EvaluatorInstance((srcDesc, dstDesc, duDesc, dvDesc, duuDesc, duvDesc, dvvDesc))

    # ---------------------------------------------------------------------------- stencils --
    @staticmethod
    def EvalStencils(srcBuffer, srcDesc, *args, instance=None, deviceContext=None, start: int = 0,
                     end: Optional[int] = None) -> bool:
        """EvalStencils(src, srcDesc, dst, dstDesc [, du, duDesc, dv, dvDesc [, duu, duuDesc, duv, duvDesc, dvv, dvvDesc]],
                        stencilTable [, instance [, deviceContext]])          (osd/cudaEvaluator.h:125-143,217-242,352-386)"""
        outs, rest = _split_outputs(args)
        if len(outs) not in (1, 3, 6) or not rest:
            raise TypeError("EvalStencils expects 1, 3 or 6 (buffer, descriptor) outputs followed by a stencil table")
        table = rest[0]
        if len(rest) > 2 and deviceContext is None:
            deviceContext = rest[2]
        n = len(outs)
        sd = _desc(srcDesc).as_c()
        dsts = (C.c_void_p * n)(*[_dev_ptr(b) for b, _ in outs])
        dds = (C.c_int * (3 * n))(*[v for _, d in outs for v in (d.offset, d.length, d.stride)])
        end = table.GetNumStencils() if end is None else end
        rc = capi.lib().b200osd_stencil_table_eval(table._h, _dev_ptr(srcBuffer), sd, n, dsts, dds, start, end,
                                                   _stream_ptr(deviceContext))
        return capi.check(rc, "B200Evaluator::EvalStencils")

    @staticmethod
    def EvalStencilsBatched(srcBuffer, srcDesc, dstBuffer, dstDesc, stencilTable, numInstances: int,
                            srcInstanceStride: int, dstInstanceStride: Optional[int] = None, deviceContext=None,
                            start: int = 0, end: Optional[int] = None) -> bool:
        """numInstances control-point sets refined through ONE table in one pass: instance b uses srcDesc.offset +
        b*srcInstanceStride / dstDesc.offset + b*dstInstanceStride (floats).  Equivalent to -- and bit-identical with --
        the reference's one-call-per-instance pattern (examples/glShareTopology/meshRefiner.h:68-88)."""
        sd, dd = _desc(srcDesc).as_c(), _desc(dstDesc).as_c()
        end = stencilTable.GetNumStencils() if end is None else end
        dstInstanceStride = srcInstanceStride if dstInstanceStride is None else dstInstanceStride
        rc = capi.lib().b200osd_stencil_table_eval_batched(stencilTable._h, _dev_ptr(srcBuffer), sd, _dev_ptr(dstBuffer), dd,
                                                           numInstances, srcInstanceStride, dstInstanceStride, start, end,
                                                           _stream_ptr(deviceContext))
        return capi.check(rc, "B200Evaluator::EvalStencilsBatched")

    @staticmethod
    def EvalStencilsRaw(src, srcDesc, outs: Sequence, sizes, offsets, indices, weights: Sequence, start: int, end: int,
                        deviceContext=None) -> bool:
        """Raw-pointer overloads on reference-layout device arrays (osd/cudaEvaluator.h:171-178,284-295,449-466).
        outs = [(dst, dstDesc), ...] (1, 3 or 6); weights = matching device weight arrays."""
        n = len(outs)
        sd = _desc(srcDesc).as_c()
        dsts = (C.c_void_p * n)(*[_dev_ptr(b) for b, _ in outs])
        dds = (C.c_int * (3 * n))(*[v for _, d in outs for v in (_desc(d).offset, _desc(d).length, _desc(d).stride)])
        ws = (C.c_void_p * n)(*[_dev_ptr(w) for w in weights[:n]])
        rc = capi.lib().b200osd_eval_stencils(_dev_ptr(src), sd, n, dsts, dds, _dev_ptr(sizes), _dev_ptr(offsets),
                                              _dev_ptr(indices), ws, start, end, _stream_ptr(deviceContext))
        return capi.check(rc, "B200Evaluator::EvalStencils(raw)")

    # ----------------------------------------------------------------------------- patches --
    @staticmethod
    def _eval_patches(srcBuffer, srcDesc, outs, numPatchCoords, patchCoords, arrays, indices, params, deviceContext,
                      options: int = 0) -> bool:
        n = len(outs)
        if n not in (1, 3, 6):
            raise TypeError("EvalPatches expects 1, 3 or 6 (buffer, descriptor) outputs")
        sd = _desc(srcDesc).as_c()
        dsts = (C.c_void_p * n)(*[_dev_ptr(b) for b, _ in outs])
        dds = (C.c_int * (3 * n))(*[v for _, d in outs for v in (d.offset, d.length, d.stride)])
        rc = capi.lib().b200osd_eval_patches_ex(_dev_ptr(srcBuffer), sd, n, dsts, dds, numPatchCoords, _dev_ptr(patchCoords),
                                                arrays, indices, params, options, _stream_ptr(deviceContext))
        return capi.check(rc, "B200Evaluator::EvalPatches")

    @staticmethod
    def _parse_patch_args(args):
        outs, rest = _split_outputs(args)
        if len(rest) < 3:
            raise TypeError("expected (numPatchCoords, patchCoords, patchTable [, fvarChannel] [, instance [, deviceContext]])")
        return outs, rest

    @staticmethod
    def _eval_patch_table(srcBuffer, srcDesc, outs, numPatchCoords, patchCoords, patchTable, which, deviceContext,
                          instance=None) -> bool:
        """Through the table handle (fast path); an instance whose bound coordinate set matches supplies
        its cached grouping, otherwise the library groups per call when that pays."""
        n = len(outs)
        if n not in (1, 3, 6):
            raise TypeError("EvalPatches expects 1, 3 or 6 (buffer, descriptor) outputs")
        sd = _desc(srcDesc).as_c()
        dsts = (C.c_void_p * n)(*[_dev_ptr(b) for b, _ in outs])
        dds = (C.c_int * (3 * n))(*[v for _, d in outs for v in (d.offset, d.length, d.stride)])
        if isinstance(instance, B200EvaluatorInstance) and instance._matches(patchTable, numPatchCoords, patchCoords):
            rc = capi.lib().b200osd_patch_plan_eval(instance._plan, which, _dev_ptr(srcBuffer), sd, n, dsts, dds, numPatchCoords,
                                                    _dev_ptr(patchCoords), _stream_ptr(deviceContext))
        else:
            rc = capi.lib().b200osd_patch_table_eval(patchTable._h, which, _dev_ptr(srcBuffer), sd, n, dsts, dds, numPatchCoords,
                                                     _dev_ptr(patchCoords), _stream_ptr(deviceContext))
        return capi.check(rc, "B200Evaluator::EvalPatches")

    @staticmethod
    def EvalPatches(srcBuffer, srcDesc, *args, deviceContext=None) -> bool:
        """EvalPatches(src, srcDesc, dst, dstDesc [, du, duDesc, dv, dvDesc [, duu.., duv.., dvv..]],
                       numPatchCoords, patchCoords, patchTable [, instance [, deviceContext]])   (osd/cudaEvaluator.h:502-677)"""
        outs, rest = B200Evaluator._parse_patch_args(args)
        n, coords, pt = rest[0], rest[1], rest[2]
        if deviceContext is None and len(rest) > 4:        # (..., patchTable, instance, deviceContext) given positionally
            deviceContext = rest[4]
        instance = rest[3] if len(rest) > 3 else None
        if isinstance(pt, B200PatchTable):
            return B200Evaluator._eval_patch_table(srcBuffer, srcDesc, outs, n, coords, pt, 0, deviceContext, instance)
        return B200Evaluator._eval_patches(srcBuffer, srcDesc, outs, n, coords, pt.GetPatchArrayBuffer(),
                                           pt.GetPatchIndexBuffer(), pt.GetPatchParamBuffer(), deviceContext)

    @staticmethod
    def EvalPatchesVarying(srcBuffer, srcDesc, *args, deviceContext=None) -> bool:
        """Same kernel on the varying triple + vertex PatchParams (osd/cudaEvaluator.h:857-1036)."""
        outs, rest = B200Evaluator._parse_patch_args(args)
        n, coords, pt = rest[0], rest[1], rest[2]
        if deviceContext is None and len(rest) > 4:
            deviceContext = rest[4]
        instance = rest[3] if len(rest) > 3 else None
        if isinstance(pt, B200PatchTable):
            return B200Evaluator._eval_patch_table(srcBuffer, srcDesc, outs, n, coords, pt, 1, deviceContext, instance)
        return B200Evaluator._eval_patches(srcBuffer, srcDesc, outs, n, coords, pt.GetVaryingPatchArrayBuffer(),
                                           pt.GetVaryingPatchIndexBuffer(), pt.GetPatchParamBuffer(), deviceContext)

    @staticmethod
    def EvalPatchesFaceVarying(srcBuffer, srcDesc, *args, deviceContext=None) -> bool:
        """Same kernel on the face-varying triple of `fvarChannel` (osd/cudaEvaluator.h:1068-1254)."""
        outs, rest = B200Evaluator._parse_patch_args(args)
        n, coords, pt = rest[0], rest[1], rest[2]
        has_ch = len(rest) > 3 and isinstance(rest[3], (int, np.integer)) and not isinstance(rest[3], bool)
        ch = int(rest[3]) if has_ch else 0
        ctx_pos = 5 if has_ch else 4                       # (..., patchTable [, fvarChannel], instance, deviceContext)
        if deviceContext is None and len(rest) > ctx_pos:
            deviceContext = rest[ctx_pos]
        instance = rest[ctx_pos - 1] if len(rest) > ctx_pos - 1 else None
        if isinstance(pt, B200PatchTable):
            return B200Evaluator._eval_patch_table(srcBuffer, srcDesc, outs, n, coords, pt, 2 + ch, deviceContext, instance)
        return B200Evaluator._eval_patches(srcBuffer, srcDesc, outs, n, coords, pt.GetFVarPatchArrayBuffer(ch),
                                           pt.GetFVarPatchIndexBuffer(ch), pt.GetFVarPatchParamBuffer(ch), deviceContext)

    @staticmethod
    def EvalPatchesRaw(src, srcDesc, outs: Sequence, numPatchCoords, patchCoords, patchArrays, patchIndices, patchParams,
                       deviceContext=None, gregory_true_derivatives: bool = False) -> bool:
        """Raw-pointer overloads (osd/cudaEvaluator.h:706-713,752-761,815-827).  gregory_true_derivatives: what the
        reference does when built with OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES (b200osd_eval_patches_ex)."""
        outs = [(b, _desc(d)) for b, d in outs]
        return B200Evaluator._eval_patches(src, srcDesc, outs, numPatchCoords, patchCoords, _dev_ptr(patchArrays),
                                           _dev_ptr(patchIndices), _dev_ptr(patchParams), deviceContext,
                                           PATCH_GREGORY_TRUE_DERIVATIVES if gregory_true_derivatives else 0)

    @staticmethod
    def Synchronize(deviceContext=None) -> None:
        capi.check(capi.lib().b200osd_synchronize(_stream_ptr(deviceContext) if deviceContext is not None else None),
                   "B200Evaluator::Synchronize")
