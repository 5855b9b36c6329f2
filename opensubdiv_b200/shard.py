"""Row-range sharding of the stencil path across the GPUs of one box (one process per GPU).

The data plane is in C (include/b200osd_capi.h: b200osd_shard_plan / _shard_coords / b200osd_comm_*, NCCL bound at run
time); this module is the Python mirror used by bench.py and the tests.  torch.distributed is only the bootstrap (it
carries the 128-byte NCCL id to the other ranks) and, in the CPU tests, the gloo stand-in for the exchange.

The reference has no multi-GPU code (no NCCL/MPI anywhere); this layer is new.  The path shards
naturally: every stencil row is independent, the tables are static, and the only per-frame shared input is the small
control-point buffer.  So

  * rows are cut into `world` contiguous ranges balanced on the number of stencil ELEMENTS (the cost of a row is its
    size, not 1) -- `balanced_row_ranges`;
  * each rank builds a B200StencilTable of its own rows only (offsets re-based, `table.row_range(a, b)`) and owns the
    matching slice of the refined buffer -- no collective ever touches tables or outputs, no reduction is needed
    because no row spans ranks;
  * once per frame the deformed control points are replicated with ONE broadcast from the rank that produced them
    (`FrameBroadcaster`, double-buffered on a side stream so frame f+1 travels while frame f is evaluated);
  * a consumer that wants the whole refined buffer on every rank calls `all_gather_rows` (optional, usually skipped);
  * EvalPatches shards by PatchCoord range (`coord_ranges` / `coord_plan`): tables replicated, every rank refines the
    broadcast control points itself and evaluates its own slice of the coordinates.

Everything here is plumbing over torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def balanced_row_ranges(sizes: np.ndarray, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """`world` contiguous row ranges [a,b) covering [0,n) with near-equal sum(sizes) (+1 per row for the fixed per-row
    cost: descriptor + output).  `align` rounds interior cut points to a multiple (e.g. the bucketing window)."""
    from . import capi
    n = int(len(sizes))
    world = max(int(world), 1)
    sz = np.ascontiguousarray(sizes, dtype=np.int32)
    out = np.zeros(2 * world, dtype=np.int32)
    rc = capi.lib().b200osd_shard_plan(n, sz.ctypes.data if n else None, world, int(align), out.ctypes.data)
    if rc != capi.OK:
        raise capi.B200OsdError("b200osd_shard_plan: " + capi.last_error())
    return [(int(out[2 * r]), int(out[2 * r + 1])) for r in range(world)]


@dataclass
class ShardPlan:
    world: int
    rank: int
    ranges: List[Tuple[int, int]]

    @property
    def start(self) -> int:
        return self.ranges[self.rank][0]

    @property
    def end(self) -> int:
        return self.ranges[self.rank][1]

    @property
    def num_rows(self) -> int:
        return self.end - self.start

    @property
    def max_rows(self) -> int:
        return max(b - a for a, b in self.ranges)

    @classmethod
    def for_table(cls, sizes: np.ndarray, world: int, rank: int, align: int = 1) -> "ShardPlan":
        return cls(world, rank, balanced_row_ranges(sizes, world, align))

    def imbalance(self, sizes: np.ndarray) -> float:
        """max shard cost / mean shard cost (1.0 = perfect)."""
        costs = [float(sizes[a:b].astype(np.int64).sum() + (b - a)) for a, b in self.ranges]
        return max(costs) / (sum(costs) / len(costs)) if sum(costs) else 1.0


@dataclass
class LocalityPlan:
    """b200osd_shard_plan_locality: rows ordered by the smallest control vertex they reference, cut into `world` chunks of
    equal cost.  `rows` = this rank's global row numbers in evaluation order; [ctrl_lo, ctrl_hi) = the control vertices
    they reference (what the rank has to receive every frame)."""
    world: int
    rank: int
    row_order: np.ndarray
    ranges: List[Tuple[int, int]]
    control_ranges: List[Tuple[int, int]]

    @classmethod
    def for_table(cls, table, world: int, rank: int) -> "LocalityPlan":
        from . import capi
        n = int(len(table.sizes))
        world = max(int(world), 1)
        sz, off, idx = (np.ascontiguousarray(x, dtype=np.int32) for x in (table.sizes, table.offsets, table.indices))
        order = np.zeros(n, dtype=np.int32)
        ranges = np.zeros(2 * world, dtype=np.int32)
        ctrl = np.zeros(2 * world, dtype=np.int32)
        p = lambda a: a.ctypes.data if a.size else None
        rc = capi.lib().b200osd_shard_plan_locality(n, p(sz), p(off), p(idx), world, p(order), ranges.ctypes.data, ctrl.ctypes.data)
        if rc != capi.OK:
            raise capi.B200OsdError("b200osd_shard_plan_locality: " + capi.last_error())
        return cls(world, rank, order, [(int(ranges[2 * r]), int(ranges[2 * r + 1])) for r in range(world)],
                   [(int(ctrl[2 * r]), int(ctrl[2 * r + 1])) for r in range(world)])

    @property
    def rows(self) -> np.ndarray:
        a, b = self.ranges[self.rank]
        return self.row_order[a:b]

    @property
    def ctrl_lo(self) -> int:
        return self.control_ranges[self.rank][0]

    @property
    def ctrl_hi(self) -> int:
        return self.control_ranges[self.rank][1]

    def imbalance(self, sizes: np.ndarray) -> float:
        costs = [float(sizes[self.row_order[a:b]].astype(np.int64).sum() + (b - a)) for a, b in self.ranges]
        return max(costs) / (sum(costs) / len(costs)) if sum(costs) else 1.0


def control_runs(table, granularity: int = 1024, max_runs: int = 8) -> List[Tuple[int, int]]:
    """b200osd_shard_control_runs: the control vertices `table` (a rank's local table) references, as index runs [lo, hi)."""
    from . import capi
    n = int(len(table.sizes))
    sz, off, idx = (np.ascontiguousarray(x, dtype=np.int32) for x in (table.sizes, table.offsets, table.indices))
    runs = np.zeros(2 * max_runs, dtype=np.int32)
    p = lambda a: a.ctypes.data if a.size else None
    k = capi.lib().b200osd_shard_control_runs(n, p(sz), p(off), p(idx), int(granularity), int(max_runs), runs.ctypes.data)
    if k < 0:
        raise capi.B200OsdError("b200osd_shard_control_runs: " + capi.last_error())
    return [(int(runs[2 * q]), int(runs[2 * q + 1])) for q in range(k)]


def local_table_rows(table, rows: np.ndarray):
    """The given rows (global row numbers, any order) as a self-contained reference-layout table, in that order."""
    from types import SimpleNamespace
    rows = np.asarray(rows, dtype=np.int64)
    sizes = np.ascontiguousarray(np.asarray(table.sizes)[rows], dtype=np.int32)
    offsets = np.zeros(len(rows), dtype=np.int32)
    if len(rows) > 1:
        np.cumsum(sizes[:-1], out=offsets[1:])
    ne = int(sizes.astype(np.int64).sum())
    # element e of the local table = element (src_off[row] + e - offsets[row]) of the global one
    src_off = np.asarray(table.offsets, dtype=np.int64)[rows]
    take = np.repeat(src_off - offsets, sizes) + np.arange(ne, dtype=np.int64)
    out = SimpleNamespace(num_control_verts=int(getattr(table, "num_control_verts", 0) or 0), sizes=sizes, offsets=offsets,
                          indices=np.ascontiguousarray(np.asarray(table.indices)[take], dtype=np.int32),
                          weights=np.ascontiguousarray(np.asarray(table.weights)[take], dtype=np.float32))
    for k in ("du", "dv", "duu", "duv", "dvv"):
        w = getattr(table, k, None)
        setattr(out, k, None if w is None else np.ascontiguousarray(np.asarray(w)[take], dtype=np.float32))
    out.num_stencils = len(rows)
    out.num_elements = ne
    return out


def coord_ranges(num_coords: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """EvalPatches shards by PatchCoord range: `world` contiguous, near-equal ranges of [0, num_coords),
    interior cuts on a multiple of `align` (a warp's worth of coordinates).  Every coordinate costs the same, so no
    weighting is needed; the patch tables are replicated (small) and each rank refines the control points it needs
    itself, so the only exchange is the same per-frame control-point broadcast the stencil path uses."""
    from . import capi
    world = max(int(world), 1)
    out = np.zeros(2 * world, dtype=np.int64)
    rc = capi.lib().b200osd_shard_coords(int(num_coords), world, int(align), out.ctypes.data)
    if rc != capi.OK:
        raise capi.B200OsdError("b200osd_shard_coords: " + capi.last_error())
    return [(int(out[2 * r]), int(out[2 * r + 1])) for r in range(world)]


def coord_plan(num_coords: int, world: int, rank: int, align: int = 32) -> ShardPlan:
    return ShardPlan(world, rank, coord_ranges(num_coords, world, align))


def local_table(table, plan: ShardPlan):
    """This rank's rows as a self-contained reference-layout table (offsets re-based)."""
    return table.row_range(plan.start, plan.end)


class FrameBroadcaster:
    """Double-buffered per-frame replication of the control points.

    `buffers` are two equally-shaped tensors (CUDA for NCCL, CPU for gloo).  `post(frame)` starts broadcasting the
    root's data for `frame` into buffers[frame % 2] on a side stream; `wait(frame)` makes the compute stream wait for
    it and returns the buffer; `release(frame)` records that the compute stream is done reading it.  With CPU
    tensors (gloo) everything degenerates to synchronous calls."""

    def __init__(self, buffers: Sequence, root: int = 0, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.buffers = list(buffers)
        assert len(self.buffers) == 2
        self.root, self.group = root, group
        self.cuda = self.buffers[0].is_cuda
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(priority=-1)
            self.ready = [torch.cuda.Event(), torch.cuda.Event()]
            self.free = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self.free:
                e.record(torch.cuda.current_stream())

    def post(self, frame: int) -> None:
        b = frame % 2
        if not self.active:
            if self.cuda:
                self.ready[b].record(self.torch.cuda.current_stream())
            return
        if self.cuda:
            self.comm_stream.wait_event(self.free[b])                 # last reader of this buffer has finished
            self.comm_stream.wait_stream(self.torch.cuda.current_stream())   # root's producer of this frame
            with self.torch.cuda.stream(self.comm_stream):
                self.dist.broadcast(self.buffers[b], src=self.root, group=self.group)
                self.ready[b].record(self.comm_stream)
        else:
            self.dist.broadcast(self.buffers[b], src=self.root, group=self.group)

    def wait(self, frame: int):
        b = frame % 2
        if self.cuda:
            self.torch.cuda.current_stream().wait_event(self.ready[b])
        return self.buffers[b]

    def release(self, frame: int) -> None:
        if self.cuda:
            self.free[frame % 2].record(self.torch.cuda.current_stream())


def all_gather_rows(local_rows, plan: ShardPlan, group=None):
    """Concatenation of every rank's refined rows (row order preserved) -- optional; shards pad to max_rows."""
    import torch
    import torch.distributed as dist
    width = local_rows.shape[1]
    pad = torch.zeros((plan.max_rows, width), dtype=local_rows.dtype, device=local_rows.device)
    pad[: plan.num_rows] = local_rows
    if not (dist.is_available() and dist.is_initialized()) or plan.world == 1:
        return pad[: plan.num_rows].clone()
    parts = [torch.empty_like(pad) for _ in range(plan.world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, plan.ranges)], dim=0)


class B200Comm:
    """The C communicator (b200osd_comm_*: NCCL over NVLink, one process per GPU).  `Create` bootstraps through an
    already initialised torch.distributed group of any backend (it only carries the 128-byte id), or takes the id."""

    def __init__(self, handle, world, rank):
        self._h, self.world, self.rank = handle, world, rank

    @staticmethod
    def available() -> bool:
        from . import capi
        return bool(capi.lib().b200osd_comm_available())

    @classmethod
    def Create(cls, world: Optional[int] = None, rank: Optional[int] = None, unique_id: Optional[bytes] = None, group=None,
               max_ctas: int = 0):
        import ctypes as C
        from . import capi
        L = capi.lib()
        if unique_id is None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            buf = (C.c_char * 128)()
            if rank == 0 and L.b200osd_comm_unique_id(buf) != capi.OK:
                raise capi.B200OsdError("b200osd_comm_unique_id: " + capi.last_error())
            box = [bytes(buf)]
            dist.broadcast_object_list(box, src=0, group=group)
            unique_id = box[0]
        idbuf = (C.c_char * 128).from_buffer_copy(unique_id)
        h = L.b200osd_comm_create_ex(int(world), int(rank), idbuf, int(max_ctas))
        if not h:
            raise capi.B200OsdError("b200osd_comm_create: " + capi.last_error())
        return cls(h, int(world), int(rank))

    def __del__(self):
        try:
            if self._h:
                from . import capi
                capi.lib().b200osd_comm_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def leak(self) -> None:
        """Skip ncclCommDestroy at interpreter exit (a communicator that recorded graphs still reference)."""
        self._h = None

    def Broadcast(self, buf, count: int, root: int = 0, deviceContext=None) -> bool:
        from . import capi
        from .osd import _dev_ptr, _stream_ptr
        return capi.check(capi.lib().b200osd_comm_broadcast(self._h, _dev_ptr(buf), int(count), int(root), _stream_ptr(deviceContext)),
                          "B200Comm::Broadcast")

    def Scatter(self, sendbuf, recvbuf, count_per_rank: int, root: int = 0, deviceContext=None) -> bool:
        from . import capi
        from .osd import _dev_ptr, _stream_ptr
        return capi.check(capi.lib().b200osd_comm_scatter(self._h, _dev_ptr(sendbuf) if sendbuf is not None else None, _dev_ptr(recvbuf),
                                                          int(count_per_rank), int(root), _stream_ptr(deviceContext)), "B200Comm::Scatter")

    def AllGather(self, sendbuf, recvbuf, count_per_rank: int, deviceContext=None) -> bool:
        from . import capi
        from .osd import _dev_ptr, _stream_ptr
        return capi.check(capi.lib().b200osd_comm_all_gather(self._h, _dev_ptr(sendbuf), _dev_ptr(recvbuf), int(count_per_rank),
                                                             _stream_ptr(deviceContext)), "B200Comm::AllGather")


class B200Window:
    """Peer-memory window (b200osd_window_*): every rank's block is addressable by every other rank; data moves by DMA,
    ordering by one-thread signal / wait kernels -- no collective kernel takes SMs from the evaluation."""

    def __init__(self, handle, comm, nbytes):
        self._h, self.comm, self.nbytes = handle, comm, nbytes
        self._tensor = None

    @classmethod
    def Create(cls, comm: B200Comm, nbytes: int) -> "B200Window":
        from . import capi
        h = capi.lib().b200osd_window_create(comm._h, int(nbytes))
        if not h:
            raise capi.B200OsdError("b200osd_window_create: " + capi.last_error())
        return cls(h, comm, int(nbytes))

    def leak(self) -> None:
        self._h = None

    def __del__(self):
        try:
            if self._h:
                from . import capi
                capi.lib().b200osd_window_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def local_tensor(self):
        """Zero-copy float32 torch view of this rank's block."""
        if self._tensor is None:
            import torch
            from . import capi
            from .osd import _CudaArrayView
            ptr = capi.lib().b200osd_window_local(self._h)
            self._tensor = torch.as_tensor(_CudaArrayView(ptr, self.nbytes // 4, self), device="cuda")
        return self._tensor

    def Get(self, src_rank: int, src_offset_bytes: int, dst, nbytes: int, deviceContext=None) -> bool:
        from . import capi
        from .osd import _dev_ptr, _stream_ptr
        return capi.check(capi.lib().b200osd_window_get(self._h, int(src_rank), int(src_offset_bytes), _dev_ptr(dst), int(nbytes),
                                                        _stream_ptr(deviceContext)), "B200Window::Get")

    def Pull(self, src_rank: int, wait_slot: int, runs, signal_rank: int = -1, signal_slot: int = -1, deviceContext=None) -> bool:
        """b200osd_window_pull: wait (wait_slot >= 0) + copy + signal (signal_slot >= 0) as one kernel.
        runs = [(src_offset_bytes, dst, nbytes), ...] (at most 8)."""
        import ctypes as C
        from . import capi
        from .osd import _dev_ptr, _stream_ptr
        k = len(runs)
        offs = (C.c_size_t * max(k, 1))(*[int(r[0]) for r in runs])
        dsts = (C.c_void_p * max(k, 1))(*[_dev_ptr(r[1]) for r in runs])
        nb = (C.c_size_t * max(k, 1))(*[int(r[2]) for r in runs])
        return capi.check(capi.lib().b200osd_window_pull(self._h, int(src_rank), int(wait_slot), k, offs, dsts, nb, int(signal_rank),
                                                         int(signal_slot), _stream_ptr(deviceContext)), "B200Window::Pull")

    def Signal(self, dst_rank: int, slot: int, deviceContext=None) -> bool:
        from . import capi
        from .osd import _stream_ptr
        return capi.check(capi.lib().b200osd_window_signal(self._h, int(dst_rank), int(slot), _stream_ptr(deviceContext)), "B200Window::Signal")

    def Wait(self, src_rank: int, slot: int, deviceContext=None) -> bool:
        from . import capi
        from .osd import _stream_ptr
        return capi.check(capi.lib().b200osd_window_wait(self._h, int(src_rank), int(slot), _stream_ptr(deviceContext)), "B200Window::Wait")

    def Error(self) -> int:
        from . import capi
        return capi.lib().b200osd_window_error(self._h)
