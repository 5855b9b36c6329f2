"""Helpers for the `-m gpu` parity tests: everything goes through the product's public API (Python mirror -> C ABI)."""
import numpy as np
import torch

import opensubdiv_b200 as osd
from opensubdiv_b200 import capi
from oracle import oracle

D = osd.BufferDescriptor


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def coords_dev(coords):
    return torch.from_numpy(np.ascontiguousarray(coords).view(np.uint8)).cuda()


def oracle_stencils(src, src_desc, n_rows, L, t, nw, start=0, end=None, abs_scale=False):
    outs = [np.zeros((n_rows, L), np.float32) for _ in range(nw)]
    ws = [t.weights, t.du, t.dv, t.duu, t.duv, t.dvv][:nw]
    def run():
        assert oracle.eval_stencils(np.ascontiguousarray(src).reshape(-1), src_desc, [o.reshape(-1) for o in outs],
                                    [(0, L, L)] * nw, t.sizes, t.offsets, t.indices, ws, start,
                                    len(t.sizes) if end is None else end)
    if abs_scale:
        with oracle.abs_mode(1):
            run()
    else:
        run()
    return outs


def oracle_patches(src, src_desc, L, coords, tr, nw, abs_scale=False, plain=False):
    outs = [np.zeros((len(coords), L), np.float32) for _ in range(nw)]
    def run():
        assert oracle.eval_patches(np.ascontiguousarray(src).reshape(-1), src_desc, [o.reshape(-1) for o in outs],
                                   [(0, L, L)] * nw, coords, tr.arrays, tr.indices, tr.params)
    if abs_scale:
        with oracle.abs_mode(1 if plain else 2):
            run()
    else:
        run()
    return outs
