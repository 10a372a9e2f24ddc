"""Short run of tools/soak_k4.py: the three K4 routes (staged decoder with production thresholds, staged decoder with tiny waves and no
direct mode, robust_kernel alone) must return identical outputs on randomized shapes and error patterns at 10^4-codeword batches."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_k4_routes_agree_on_random_batches():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak_k4.py"), "--seconds", "10", "--seed", "3"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "soak ok" in res.stdout
