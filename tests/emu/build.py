"""Builds tests/emu/libpatch_emu.so: the CUDA patch kernel's per-coordinate code compiled for the HOST (test-only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpatch_emu.so")
SRC = os.path.join(HERE, "patch_emu.cu")
CSRC = os.path.join(HERE, "..", "..", "opensubdiv_b200", "csrc")
DEPS = [os.path.join(CSRC, f) for f in ("patch_kernels.cuh", "patchmap.cuh", "common.cuh")]


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) > max(os.path.getmtime(f) for f in [SRC] + DEPS):
        return LIB
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
                           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, SRC],
                          stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    return LIB


if __name__ == "__main__":
    print(build(True))
