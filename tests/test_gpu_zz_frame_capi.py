"""C-ABI frame capture (b200osd_frame_*) on a real B200.  The check runs in a subprocess (tests/
frame_capi_check.py) so that a failure cannot disturb the rest of the GPU suite."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.gpu
def test_frame_capture_through_the_c_abi():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "frame_capi_check.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=180)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and "FRAME CAPI OK" in r.stdout
