"""Builds opensubdiv_b200/libb200osd.so (the C-ABI library) with nvcc for sm_100a, in-tree.

    python -m opensubdiv_b200._build [--force]

nvcc cross-compiles without a GPU.  The result is git-ignored but travels to the GPU box with the
gpurun snapshot, so nothing is compiled there.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libb200osd.so")
SOURCES = ["core.cu", "stencil.cu", "patch.cu", "patchmap.cu", "frame.cu", "shard.cu"]
HEADERS = ["common.cuh", "stencil_kernels.cuh", "patch_kernels.cuh", "patchmap.cuh", os.path.join("..", "..", "include", "b200osd_capi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # no implicit mul+add contraction: every FMA in the kernels is written as fmaf(), so the arithmetic is exactly what the
    # source says, identical across kernels that share device functions (caller-order / hull-cache / grouped patch paths
    # must agree bit for bit) and identical to the host emulation in tests/emu (-ffp-contract=off)
    "-fmad=false",
    "-Xptxas=-v", "--threads", "4",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "-shared", "-ldl",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # /opt/gcc (the image's default $CXX) is fine for nvcc's host pass, but keep it deterministic:
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    log = os.path.join(_HERE, "csrc", "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed (see %s)" % log)
    if verbose:
        sys.stdout.write(r.stdout)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB_PATH)
