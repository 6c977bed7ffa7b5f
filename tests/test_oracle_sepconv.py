"""Pin the Sepconv oracle (oracle/sepconv.py) to outputs of the unmodified reference Network (tests/golden/sepconv_*.npz,
made by tools/make_golden_sepconv.py with the custom op replaced by its CPU restatement)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_sepconv import sepconv_cases, sepconv_inputs  # noqa: E402
from oracle import sepconv as OS  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name", sorted(sepconv_cases().keys()))
def test_sepconv_oracle_matches_reference_output(name):
    cfg = sepconv_cases()[name]
    ref = torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["out"])
    sd = OS.synthetic_state_dict(cfg["seed"])
    x = sepconv_inputs(cfg).permute(0, 3, 1, 2).contiguous()
    out = OS.network_forward(sd, x[0:1], x[1:2])
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-5
    assert float(ref.min()) > -0.2 and float(ref.max()) < 1.2   # the synthetic kernels are well conditioned


def test_sepconv_state_dict_spec():
    spec = OS.state_dict_spec()
    assert len(spec) == 88
    assert sum(int(np.prod(s)) for _, s in spec) == 13560102
    d = dict(spec)
    assert d["netDecode.0.netVer.1.netMain.2.weight"] == (256, 512, 3, 3)
    assert d["netHortwo.netMain.3.weight"] == (51, 64, 3, 3)
