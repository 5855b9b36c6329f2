"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm prints ONE JSON line
with the agreed keys on the same `config` the B200 arm uses, and the B200 arm refuses to run without a device (no CPU
fallback).  The B200 arm itself is run on the GPU box by the driver."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libosdref.so not built (needs /root/reference)")
def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "refined_verts_per_sec_EvalStencils" and d["unit"] == "verts/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3
    assert d["value"] > 1e6 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    cfg = d["config"]
    assert cfg["workload"].startswith("catmark") and cfg["rows"] == 6_400_000 and cfg["primvar_floats"] == 6
    assert "far_insertion_order" in cfg["table_order"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == pytest.approx(d["value"]) and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == pytest.approx(d["value"]) and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # the same `config` as the B200 arm builds (bench.shared_config is the single source of both)
    sys.path.insert(0, ROOT)
    import bench
    assert set(cfg) == {"workload", "rows", "elements", "control_verts", "primvar_floats", "table_order", "l2_policy"}
    assert "larger than L2" in cfg["l2_policy"]
    assert callable(bench.shared_config)


@pytest.mark.skipif(torch.cuda.is_available(), reason="a device is present")
def test_b200_arm_refuses_to_run_without_a_device():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
