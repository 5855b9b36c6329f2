"""opensubdiv_b200 -- B200-native (sm_100a) Osd evaluator: EvalStencils / EvalPatches* behind the Osd API.

The product is libb200osd.so (hand-written CUDA + a C ABI, include/b200osd_capi.h).  This package is the
Python host-side mirror of the reference Osd interface plus the multi-GPU sharding helpers.
"""
from .osd import (BufferDescriptor, B200VertexBuffer, B200StencilTable, B200PatchTable, B200PatchMap, B200FrameGraph, B200Evaluator,
                  PATCH_COORD_DTYPE, PATCH_ARRAY_DTYPE, PATCH_PARAM_DTYPE)
from .capi import B200OsdError

__all__ = ["BufferDescriptor", "B200VertexBuffer", "B200StencilTable", "B200PatchTable", "B200PatchMap", "B200FrameGraph", "B200Evaluator",
           "B200OsdError", "PATCH_COORD_DTYPE", "PATCH_ARRAY_DTYPE", "PATCH_PARAM_DTYPE"]
