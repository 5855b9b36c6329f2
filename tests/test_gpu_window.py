"""The peer-memory window on ONE GPU (world 1: the window's source is the rank itself): b200osd_window_get and the fused
b200osd_window_pull kernel copy byte ranges correctly, including unaligned pieces and 4-byte tails.  The multi-rank
ordering (signal / wait) is exercised by bench.py --gpus N (weak scaling and the config-5 section)."""
import ctypes as C

import numpy as np
import pytest
import torch

from opensubdiv_b200 import capi, shard

pytestmark = pytest.mark.gpu


def test_window_pull_copies_runs_like_get():
    if not shard.B200Comm.available():
        pytest.skip("NCCL not loadable")
    L = capi.lib()
    ident = (C.c_char * 128)()
    assert L.b200osd_comm_unique_id(ident) == capi.OK
    comm = shard.B200Comm.Create(1, 0, bytes(ident))
    n = 300_000
    win = shard.B200Window.Create(comm, n * 4)
    src = win.local_tensor()
    src.copy_(torch.arange(n, dtype=torch.float32, device="cuda") * 0.5)
    # (source offset in floats, floats): aligned, odd-sized, 4-byte-aligned only, tiny
    runs = [(0, 65536), (70000, 12345), (100001, 40003), (299990, 10)]
    a = torch.full((n,), -1.0, device="cuda")
    b = torch.full((n,), -1.0, device="cuda")
    pulls = [(off * 4, a[off:off + cnt], cnt * 4) for off, cnt in runs]
    assert win.Pull(0, -1, pulls, -1, -1, None)
    for off, cnt in runs:
        assert win.Get(0, off * 4, b[off:off + cnt], cnt * 4, None)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    want = np.full(n, -1.0, np.float32)
    for off, cnt in runs:
        want[off:off + cnt] = np.arange(off, off + cnt, dtype=np.float32) * 0.5
    assert np.array_equal(a.cpu().numpy(), want)
    assert win.Error() == 0
    # argument checks
    assert not win.Pull(0, -1, [(n * 4 - 8, a, 16)], -1, -1, None)            # outside the window
    assert not win.Pull(0, -1, [(0, a, 6)], -1, -1, None)                     # not a multiple of 4 bytes
    assert not win.Pull(0, -1, [(0, a, 4)] * 9, -1, -1, None)                 # more than 8 runs
    # wait / signal slots on a world of one are no-ops, not hangs
    assert win.Pull(0, 3, pulls[:1], -1, 4, None)
    torch.cuda.synchronize()
    assert win.Error() == 0
