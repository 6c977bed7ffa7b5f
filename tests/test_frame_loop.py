"""generic_frame_loop (host logic) vs the reference's own outputs (tests/golden/loop_*.npz from tools/make_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import loop_cases, loop_frames, loop_model  # noqa: E402


@pytest.mark.parametrize("name", sorted(loop_cases().keys()))
def test_loop_matches_reference(pkg, name):
    from cfi_b200.frame_loop import generic_frame_loop
    from cfi_b200.node import InterpolationStateList
    cfg = loop_cases()[name]
    ref = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))["out"])
    st = None if cfg["states"] is None else InterpolationStateList(list(cfg["states"][0]), cfg["states"][1])
    out = generic_frame_loop("Stand_In_VFI", loop_frames(cfg["n"]), 10, cfg["multiplier"], loop_model, 0.03,
                             interpolation_states=st, use_timestep=cfg["use_timestep"], dtype=torch.float32)
    assert out.shape == ref.shape and not out.is_cuda
    assert (out - ref).abs().max().item() < 1e-6


def test_loop_needs_two_frames(pkg):
    from cfi_b200.frame_loop import generic_frame_loop
    with pytest.raises(AssertionError):
        generic_frame_loop("X_VFI", torch.zeros(1, 3, 4, 4), 10, 2, loop_model, 0.0)
