"""Sepconv on the GPU: tools/sepconv_gpu_check.py in a SUBPROCESS (sorted after the verified
tests; a trap in new code must not poison their CUDA context).  The Sepconv trunk was written at the end of r01 with no
GPU minutes left - streamconv's EXT epilogue, nine element-wise kernels and the schedule have never run on a GPU - so the
test is xfail(strict=False): it still runs on the GPU box at round end and reports XPASS / XFAIL with the checker's
per-stage JSON lines."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sepconv_gpu_check_subprocess():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sepconv_gpu_check.py")],
                       capture_output=True, text=True, timeout=420, cwd=ROOT)
    tail = (r.stdout[-6000:] + "\n--- stderr ---\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
