"""CPU-side checks of the product's host layer: the C-ABI library loads and exports every declared symbol, argument
validation mirrors the reference, and nothing silently falls back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import opensubdiv_b200 as osd
from opensubdiv_b200 import capi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200osd_capi.h")).read()
    declared = sorted(set(re.findall(r"\b(b200osd_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    lib = capi.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == declared


def test_pod_layouts_match_reference_sizes():
    assert osd.PATCH_COORD_DTYPE.itemsize == 20      # osd/types.h:42-64
    assert osd.PATCH_ARRAY_DTYPE.itemsize == 24      # osd/types.h:66-122
    assert osd.PATCH_PARAM_DTYPE.itemsize == 12      # osd/types.h:127-130


def test_buffer_descriptor_semantics():
    d = osd.BufferDescriptor(16, 3, 13)              # osd/bufferDescriptor.h:71-78
    assert d.GetLocalOffset() == 3 and d.IsValid()
    assert not osd.BufferDescriptor(11, 3, 13).IsValid()
    assert not osd.BufferDescriptor().IsValid()


def test_validation_happens_before_any_device_work():
    """These return the reference's `false` (ERR_INVALID) without touching CUDA, so they are checkable on CPU."""
    lib = capi.lib()
    sd = (C.c_int * 3)(0, 3, 3)
    dd = (C.c_int * 3)(0, 4, 4)
    dsts = (C.c_void_p * 1)(1 << 20)
    w = (C.c_void_p * 1)(1 << 20)
    # length mismatch -> false (osd/cpuEvaluator.cpp:47)
    assert lib.b200osd_eval_stencils(1 << 20, sd, 1, dsts, dd, 1 << 20, 1 << 20, 1 << 20, w, 0, 10, None) == capi.ERR_INVALID
    # end <= start -> true, no-op (osd/cpuEvaluator.cpp:46)
    assert lib.b200osd_eval_stencils(1 << 20, sd, 1, dsts, dd, 1 << 20, 1 << 20, 1 << 20, w, 7, 7, None) == capi.OK
    # NULL dst in the value-only form -> false (osd/cudaEvaluator.cpp:159)
    nul = (C.c_void_p * 1)(None)
    dd3 = (C.c_int * 3)(0, 3, 3)
    assert lib.b200osd_eval_stencils(1 << 20, sd, 1, nul, dd3, 1 << 20, 1 << 20, 1 << 20, w, 0, 10, None) == capi.ERR_INVALID
    # EvalPatches: NULL src -> false (osd/cpuEvaluator.cpp:165-169)
    assert lib.b200osd_eval_patches(None, sd, 1, dsts, dd3, 5, 1 << 20, 1 << 20, 1 << 20, 1 << 20, None) == capi.ERR_INVALID
    assert lib.b200osd_eval_patches(1 << 20, sd, 2, dsts, dd3, 5, 1 << 20, 1 << 20, 1 << 20, 1 << 20, None) == capi.ERR_INVALID
    assert b"nOut" in lib.b200osd_last_error()


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert osd.B200VertexBuffer.Create(3, 16) is None          # Create() -> NULL like the reference, never a host buffer
    assert "cuda" in capi.last_error().lower()
    t = type("T", (), dict(sizes=np.array([1], np.int32), offsets=np.array([0], np.int32),
                           indices=np.array([0], np.int32), weights=np.array([1.0], np.float32)))()
    assert osd.B200StencilTable.Create(t) is None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opensubdiv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower().replace("no reference code involved", ""), f


def test_header_is_plain_c_and_cxx():
    """The boundary is a C ABI: the header must compile as C99 and as C++ without anything else on the include path."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "b200osd_capi.h")
    if not shutil.which("gcc"):
        pytest.skip("gcc not found")
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):      # long long: C99 / C++11
        subprocess.check_call(["gcc", "-x", lang, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", hdr])


def test_patch_map_validation_without_a_gpu():
    lib = capi.lib()
    from tests.util import golden
    d = golden("patchmap_catmark_cube")
    arrays = np.ascontiguousarray(d["arrays"]).copy()
    params = np.ascontiguousarray(d["params"])
    # inconsistent table: rejected on the host before any device work, with a message
    bad = arrays.copy()
    bad["primitiveIdBase"][0] = 5
    assert not lib.b200osd_patch_map_create(len(bad), bad.ctypes.data, len(params), params.ctypes.data, 0)
    assert b"tile" in lib.b200osd_last_error()
    assert lib.b200osd_patch_map_find(None, 4, None, 1, None, 1, None, 1, None, None, None) == capi.ERR_INVALID
    import torch
    if not torch.cuda.is_available():
        # a valid table still cannot be created without a device: no host-side stand-in
        assert not lib.b200osd_patch_map_create(len(arrays), arrays.ctypes.data, len(params), params.ctypes.data, 0)
        with pytest.raises(osd.B200OsdError):
            osd.B200PatchMap.Create(type("PT", (), dict(vertex=type("T", (), dict(arrays=arrays, params=params))(),
                                                        varying=None, fvar=[]))(), patchesAreTriangular=False)


def test_positional_device_context_reaches_the_c_abi(monkeypatch):
    """EvalPatches*(..., patchTable [, fvarChannel], instance, deviceContext): the trailing positional deviceContext of
    the reference signatures (osd/cudaEvaluator.h:502-523, 1068-1090) must select the stream, like the keyword form."""
    seen = {}

    def fake(src, desc, outs, n, coords, pt, which, ctx, instance=None):
        seen["which"], seen["ctx"] = which, ctx
        seen_instance.append(instance)
        return True
    seen_instance = []
    monkeypatch.setattr(osd.B200Evaluator, "_eval_patch_table", staticmethod(fake))
    pt = osd.B200PatchTable(None)
    D = osd.BufferDescriptor
    assert osd.B200Evaluator.EvalPatches(1, D(0, 3, 3), 2, D(0, 3, 3), 10, 3, pt, None, 77)
    assert seen == {"which": 0, "ctx": 77}
    assert osd.B200Evaluator.EvalPatches(1, D(0, 3, 3), 2, D(0, 3, 3), 10, 3, pt, None)
    assert seen == {"which": 0, "ctx": None}
    assert osd.B200Evaluator.EvalPatchesVarying(1, D(0, 3, 3), 2, D(0, 3, 3), 10, 3, pt, None, deviceContext=5)
    assert seen == {"which": 1, "ctx": 5}
    assert osd.B200Evaluator.EvalPatchesFaceVarying(1, D(0, 2, 2), 2, D(0, 2, 2), 10, 3, pt, 1, None, 9)
    assert seen == {"which": 3, "ctx": 9}
    assert osd.B200Evaluator.EvalPatchesFaceVarying(1, D(0, 2, 2), 2, D(0, 2, 2), 10, 3, pt, 0, None)
    assert seen == {"which": 2, "ctx": None}
    assert osd.B200Evaluator.EvalPatchesFaceVarying(1, D(0, 2, 2), 2, D(0, 2, 2), 10, 3, pt)
    assert seen == {"which": 2, "ctx": None}
    assert all(i is None for i in seen_instance)
    # the instance of the "instantiatable" flavour (osd/mesh.h:305-409) travels in the reference's positional slot
    inst = osd.B200Evaluator.Create(D(0, 3, 3), D(0, 3, 3))
    assert osd.B200Evaluator.EvalPatches(1, D(0, 3, 3), 2, D(0, 3, 3), 10, 3, pt, inst, 4)
    assert seen == {"which": 0, "ctx": 4} and seen_instance[-1] is inst
    assert osd.B200Evaluator.EvalPatchesFaceVarying(1, D(0, 2, 2), 2, D(0, 2, 2), 10, 3, pt, 1, inst)
    assert seen == {"which": 3, "ctx": None} and seen_instance[-1] is inst


def test_ctypes_prototypes_match_the_header_arity():
    """Every declared entry point that takes arguments must have ctypes argtypes of the same arity (a missing prototype
    would pass 64-bit pointers as C ints)."""
    header = open(os.path.join(ROOT, "include", "b200osd_capi.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    lib = capi.lib()
    checked = 0
    for m in re.finditer(r"\b(b200osd_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        name, params = m.group(1), " ".join(m.group(2).split())
        arity = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(lib, name)
        if arity == 0:
            continue
        assert fn.argtypes is not None, f"{name}: no ctypes prototype"
        assert len(fn.argtypes) == arity, f"{name}: header declares {arity} parameters, ctypes has {len(fn.argtypes)}"
        checked += 1
    assert checked >= 30
    # ... and every entry point returning a pointer or a long long must declare it (default restype is a C int)
    for m in re.finditer(r"B200OSD_API\s+([^;(]*?)\b(b200osd_[a-z_0-9]+)\s*\(", header):
        rtype, name = " ".join(m.group(1).split()), m.group(2)
        if "*" in rtype or "long long" in rtype:
            assert getattr(lib, name).restype is not C.c_int, f"{name} returns '{rtype}' but ctypes restype is int"


def test_plain_c_example_builds_against_the_abi(tmp_path):
    """examples/c_abi_example.c: a C99 client of the library (compiles with -pedantic, links, and without a device fails
    loudly instead of computing on the host)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not found")
    exe = str(tmp_path / "c_abi_example")
    libdir = os.path.join(ROOT, "opensubdiv_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", libdir, "-lb200osd",
                           "-Wl,-rpath," + libdir, "-o", exe])
    import torch
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "refined[4] = 0.5 0.5 0.125" in r.stdout, r.stdout
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stdout


def test_plain_c_sharding_example_runs_on_the_host(tmp_path):
    """examples/c_shard_example.c: the partitioning entries of the multi-GPU data plane (b200osd_shard_plan,
    _shard_plan_locality, _shard_control_runs) called from C99 -- host-only functions, so the example runs here."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not found")
    exe = str(tmp_path / "c_shard_example")
    libdir = os.path.join(ROOT, "opensubdiv_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_shard_example.c"), "-L", libdir, "-lb200osd",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    assert "rank 0: contiguous rows [0,16) | by locality 16 rows, control vertices [0,64) in 2 run(s): [0,16) [63,64) = 17 of 64" in r.stdout
    assert "rank 3:" in r.stdout and "[47,64) = 17 of 64" in r.stdout
