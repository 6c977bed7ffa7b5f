"""Sepconv on the GPU: tools/sepconv_gpu_check.py in a SUBPROCESS (sorted after the verified
tests; a trap in new code must not poison their CUDA context).  The Sepconv trunk was written at the end of r01 with no
GPU minutes left; its first GPU run (r02, profiles/r02_sepconv_gpu_check.jsonl) was green on every stage: 98 - 101 dB on
the tcgen05 kernel vs the unmodified reference Network."""
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
