"""FILM on the GPU.  Runs tools/film_gpu_check.py in a SUBPROCESS - the FILM kernels
were written in round 1 without GPU access, so a kernel trap must not poison the CUDA context of the (GPU-verified)
RIFE tests that share this pytest process; the file name sorts it after them.  Marked xfail(strict=False) until its
first GPU session has been read: the run still executes on the GPU box and reports XPASS / XFAIL with the per-stage
JSON lines of the checker in the failure text."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(reason="FILM CUDA path not yet run on a GPU (written in r01 after the GPU budget was spent)",
                   strict=False)
def test_film_gpu_check_subprocess():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "film_gpu_check.py"), "--quick"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout[-6000:] + "\n--- stderr ---\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
