"""C-ABI frame capture (b200osd_frame_*) on a real B200.  The check runs in a subprocess (tests/
frame_capi_check.py).  This file was written after the round's GPU budget was spent, so it has not run on hardware
yet: it is non-strict xfail until it has (the torch-driven capture of the same calls IS verified,
tests/test_gpu_frame_graph.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="not yet run on hardware (written after the round's GPU budget was spent)")
def test_frame_capture_through_the_c_abi():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "frame_capi_check.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=180)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and "FRAME CAPI OK" in r.stdout
