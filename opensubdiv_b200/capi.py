"""ctypes binding of libb200osd.so -- the C ABI declared in include/b200osd_capi.h.

There is no CPU fallback: if the library has not been built this module raises at load time, and
every compute entry point returns an error code when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200osd.so")

OK, ERR_INVALID, ERR_CUDA, ERR_ALLOC, ERR_UNSUPPORTED = 0, 1, 2, 3, 4

# every symbol include/b200osd_capi.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "b200osd_version", "b200osd_last_error", "b200osd_launch_count", "b200osd_reset_launch_count",
    "b200osd_synchronize",
    "b200osd_vertex_buffer_create", "b200osd_vertex_buffer_destroy", "b200osd_vertex_buffer_num_elements",
    "b200osd_vertex_buffer_num_vertices", "b200osd_vertex_buffer_bind", "b200osd_vertex_buffer_update",
    "b200osd_vertex_buffer_read",
    "b200osd_stencil_table_create", "b200osd_stencil_table_create_from_device", "b200osd_stencil_table_destroy", "b200osd_stencil_table_num_stencils",
    "b200osd_stencil_table_num_control_vertices", "b200osd_stencil_table_num_elements",
    "b200osd_stencil_table_is_factorized",
    "b200osd_stencil_table_buffer", "b200osd_stencil_table_stream_bytes", "b200osd_stencil_table_eval",
    "b200osd_stencil_table_eval_batched", "b200osd_stencil_table_set_variant", "b200osd_stencil_table_get_variant",
    "b200osd_eval_stencils", "b200osd_limit_stencil_table_create",
    "b200osd_patch_table_create", "b200osd_patch_table_destroy", "b200osd_patch_table_set",
    "b200osd_patch_table_num_fvar_channels", "b200osd_patch_table_buffer", "b200osd_patch_table_count",
    "b200osd_eval_patches", "b200osd_eval_patches_ex", "b200osd_patch_table_eval", "b200osd_patch_table_set_variant",
    "b200osd_patch_table_get_variant", "b200osd_patch_table_set_options", "b200osd_patch_table_get_options",
    "b200osd_patch_plan_create", "b200osd_patch_plan_destroy", "b200osd_patch_plan_capacity", "b200osd_patch_plan_bin",
    "b200osd_patch_plan_eval",
    "b200osd_patch_map_create", "b200osd_patch_map_destroy", "b200osd_patch_map_info", "b200osd_patch_map_find",
    "b200osd_frame_side_stream", "b200osd_frame_fence", "b200osd_frame_set_l2_window",
    "b200osd_shard_plan", "b200osd_shard_plan_locality", "b200osd_shard_control_runs", "b200osd_shard_coords", "b200osd_comm_available", "b200osd_comm_unique_id", "b200osd_comm_create", "b200osd_comm_create_ex",
    "b200osd_comm_destroy", "b200osd_comm_world", "b200osd_comm_rank", "b200osd_comm_broadcast", "b200osd_comm_scatter",
    "b200osd_comm_all_gather",
    "b200osd_window_create", "b200osd_window_destroy", "b200osd_window_local", "b200osd_window_bytes", "b200osd_window_get", "b200osd_window_pull",
    "b200osd_window_signal", "b200osd_window_wait", "b200osd_window_error",
    "b200osd_frame_create", "b200osd_frame_destroy", "b200osd_frame_stream", "b200osd_frame_begin", "b200osd_frame_end",
    "b200osd_frame_launch", "b200osd_frame_synchronize",
]


class B200OsdError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200OsdError(
            f"{LIB_PATH} is missing: build it with `python -m opensubdiv_b200._build` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
    L.b200osd_version.restype = C.c_char_p
    L.b200osd_last_error.restype = C.c_char_p
    L.b200osd_launch_count.restype = ll
    L.b200osd_synchronize.argtypes = [vp]
    L.b200osd_vertex_buffer_create.restype = vp
    L.b200osd_vertex_buffer_create.argtypes = [i, i]
    L.b200osd_vertex_buffer_destroy.argtypes = [vp]
    L.b200osd_vertex_buffer_num_elements.argtypes = [vp]
    L.b200osd_vertex_buffer_num_vertices.argtypes = [vp]
    L.b200osd_vertex_buffer_bind.restype = vp
    L.b200osd_vertex_buffer_bind.argtypes = [vp]
    L.b200osd_vertex_buffer_update.argtypes = [vp, vp, i, i, vp]
    L.b200osd_vertex_buffer_read.argtypes = [vp, vp, i, i, vp]
    L.b200osd_stencil_table_create.restype = vp
    L.b200osd_stencil_table_create.argtypes = [i, i] + [vp] * 9 + [i]
    L.b200osd_stencil_table_create_from_device.restype = vp
    L.b200osd_stencil_table_create_from_device.argtypes = [i, i] + [vp] * 9 + [i]
    L.b200osd_stencil_table_destroy.argtypes = [vp]
    L.b200osd_stencil_table_num_stencils.argtypes = [vp]
    L.b200osd_stencil_table_num_control_vertices.argtypes = [vp]
    L.b200osd_stencil_table_num_elements.argtypes = [vp]
    L.b200osd_stencil_table_num_elements.restype = ll
    L.b200osd_stencil_table_buffer.restype = vp
    L.b200osd_stencil_table_buffer.argtypes = [vp, i]
    L.b200osd_stencil_table_stream_bytes.restype = ll
    L.b200osd_stencil_table_stream_bytes.argtypes = [vp, i]
    L.b200osd_stencil_table_eval.argtypes = [vp, vp, vp, i, vp, vp, i, i, vp]
    L.b200osd_stencil_table_eval_batched.argtypes = [vp, vp, vp, vp, vp, i, ll, ll, i, i, vp]
    L.b200osd_eval_stencils.argtypes = [vp, vp, i, vp, vp, vp, vp, vp, vp, i, i, vp]
    L.b200osd_limit_stencil_table_create.restype = vp
    L.b200osd_limit_stencil_table_create.argtypes = [vp, vp, i, vp, i, i, vp]
    L.b200osd_patch_table_create.restype = vp
    L.b200osd_patch_table_create.argtypes = [i]
    L.b200osd_patch_table_destroy.argtypes = [vp]
    L.b200osd_patch_table_set.argtypes = [vp, i, i, vp, i, vp, i, vp]
    L.b200osd_patch_table_num_fvar_channels.argtypes = [vp]
    L.b200osd_patch_table_buffer.restype = vp
    L.b200osd_patch_table_buffer.argtypes = [vp, i, i]
    L.b200osd_patch_table_count.argtypes = [vp, i, i]
    L.b200osd_eval_patches.argtypes = [vp, vp, i, vp, vp, i, vp, vp, vp, vp, vp]
    L.b200osd_eval_patches_ex.argtypes = [vp, vp, i, vp, vp, i, vp, vp, vp, vp, i, vp]
    L.b200osd_patch_table_eval.argtypes = [vp, i, vp, vp, i, vp, vp, i, vp, vp]
    L.b200osd_patch_table_set_options.argtypes = [vp, i]
    L.b200osd_patch_table_get_options.argtypes = [vp]
    L.b200osd_patch_table_set_variant.argtypes = [vp, i]
    L.b200osd_patch_table_get_variant.argtypes = [vp]
    L.b200osd_patch_plan_create.restype = vp
    L.b200osd_patch_plan_create.argtypes = [vp, i]
    L.b200osd_patch_plan_destroy.argtypes = [vp]
    L.b200osd_patch_plan_capacity.argtypes = [vp]
    L.b200osd_patch_plan_bin.argtypes = [vp, i, vp, vp]
    L.b200osd_patch_plan_eval.argtypes = [vp, i, vp, vp, i, vp, vp, i, vp, vp]
    L.b200osd_patch_map_create.restype = vp
    L.b200osd_patch_map_create.argtypes = [i, vp, i, vp, i]
    L.b200osd_patch_map_destroy.argtypes = [vp]
    L.b200osd_patch_map_info.argtypes = [vp, vp]
    L.b200osd_patch_map_find.argtypes = [vp, i, vp, i, vp, i, vp, i, vp, vp, vp]
    L.b200osd_stencil_table_set_variant.argtypes = [vp, i]
    L.b200osd_stencil_table_get_variant.argtypes = [vp]
    L.b200osd_stencil_table_is_factorized.argtypes = [vp]
    L.b200osd_frame_create.restype = vp
    L.b200osd_frame_destroy.argtypes = [vp]
    L.b200osd_frame_stream.restype = vp
    L.b200osd_frame_stream.argtypes = [vp]
    for fn in (L.b200osd_frame_begin, L.b200osd_frame_end, L.b200osd_frame_launch, L.b200osd_frame_synchronize):
        fn.argtypes = [vp]
    L.b200osd_frame_side_stream.restype = vp
    L.b200osd_frame_side_stream.argtypes = [vp]
    L.b200osd_frame_fence.argtypes = [vp, i]
    L.b200osd_frame_set_l2_window.argtypes = [vp, vp, C.c_size_t, C.c_float]
    L.b200osd_shard_plan.argtypes = [i, vp, i, i, vp]
    L.b200osd_shard_plan_locality.argtypes = [i, vp, vp, vp, i, vp, vp, vp]
    L.b200osd_shard_control_runs.argtypes = [i, vp, vp, vp, i, i, vp]
    L.b200osd_shard_coords.argtypes = [ll, i, i, vp]
    L.b200osd_comm_unique_id.argtypes = [vp]
    L.b200osd_comm_create.restype = vp
    L.b200osd_comm_create.argtypes = [i, i, vp]
    L.b200osd_comm_create_ex.restype = vp
    L.b200osd_comm_create_ex.argtypes = [i, i, vp, i]
    L.b200osd_comm_destroy.argtypes = [vp]
    L.b200osd_comm_world.argtypes = [vp]
    L.b200osd_comm_rank.argtypes = [vp]
    L.b200osd_comm_broadcast.argtypes = [vp, vp, C.c_size_t, i, vp]
    L.b200osd_comm_scatter.argtypes = [vp, vp, vp, C.c_size_t, i, vp]
    L.b200osd_comm_all_gather.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.b200osd_window_create.restype = vp
    L.b200osd_window_create.argtypes = [vp, C.c_size_t]
    L.b200osd_window_destroy.argtypes = [vp]
    L.b200osd_window_local.restype = vp
    L.b200osd_window_local.argtypes = [vp]
    L.b200osd_window_bytes.restype = C.c_size_t
    L.b200osd_window_bytes.argtypes = [vp]
    L.b200osd_window_get.argtypes = [vp, i, C.c_size_t, vp, C.c_size_t, vp]
    L.b200osd_window_pull.argtypes = [vp, i, i, i, vp, vp, vp, i, i, vp]
    L.b200osd_window_signal.argtypes = [vp, i, i, vp]
    L.b200osd_window_wait.argtypes = [vp, i, i, vp]
    L.b200osd_window_error.argtypes = [vp]
    _lib = L
    return L


def last_error() -> str:
    return lib().b200osd_last_error().decode(errors="replace")


def check(rc: int, what: str) -> bool:
    """OK -> True; ERR_INVALID -> False (the reference evaluators' `return false`); anything else raises."""
    if rc == OK:
        return True
    if rc == ERR_INVALID:
        return False
    raise B200OsdError(f"{what} failed (code {rc}): {last_error()}")
