"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import os
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: results must match Osd::CpuEvaluator within 1e-6 relative error in fp32.  "Relative" is taken against
# max(|ref|, S) with S = sum_j |w_j||x_j| (the oracle's abs mode): derivative weights are signed and cancel, so an
# element-wise relative bound against |ref| alone is meaningless near zeros.  Differences come
# only from summation order (separable B-spline evaluation) and FMA contraction on the GPU.
REL_TOL = 1e-6


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def table_from(d, prefix):
    """Rebuilds a stencil-table namespace (sizes, offsets, indices, weights, du..dvv) from a fixture."""
    t = SimpleNamespace(num_control_verts=int(d[prefix + "ncv"]), sizes=d[prefix + "sizes"], offsets=d[prefix + "offsets"],
                        indices=d[prefix + "indices"], weights=d[prefix + "weights"])
    for k in ("du", "dv", "duu", "duv", "dvv"):
        setattr(t, k, d[prefix + k] if (prefix + k) in d.files else None)
    t.num_stencils = len(t.sizes)
    return t


def triple_from(d, prefix):
    return SimpleNamespace(arrays=d[prefix + "arrays"], indices=d[prefix + "indices"], params=d[prefix + "params"])


def weight_streams(t, nw):
    return [t.weights, t.du, t.dv, t.duu, t.duv, t.dvv][:nw]


def assert_close(got, ref, scale, what="", tol=REL_TOL):
    got, ref, scale = np.asarray(got, np.float64), np.asarray(ref, np.float64), np.asarray(scale, np.float64)
    denom = np.maximum(np.abs(ref), np.abs(scale))
    denom = np.maximum(denom, 1e-30)
    err = np.abs(got - ref) / denom
    worst = float(err.max()) if err.size else 0.0
    assert worst <= tol, f"{what}: max relative error {worst:.3e} > {tol:.1e} at {np.unravel_index(err.argmax(), err.shape)}"
    return worst


def assert_close_bbox(got, ref, bbox, depth, order, what="", tol=REL_TOL):
    """Second, scale-free gate for patch evaluation: |got - ref| <= tol * bbox * 2^(order * depth), with
    bbox the extent of the control points and depth the patch's subdivision depth -- a derivative of order k of a
    depth-d sub-patch is naturally 2^(k d) times larger than the geometry.  Returns the worst ratio error / (bbox 2^(k d))."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = float(bbox) * np.power(2.0, order * np.asarray(depth, np.float64))[:, None]
    err = np.abs(got - ref) / scale
    worst = float(err.max()) if err.size else 0.0
    assert worst <= tol, f"{what}: error {worst:.3e} of bbox*2^(order*depth) > {tol:.1e} at {np.unravel_index(err.argmax(), err.shape)}"
    return worst
