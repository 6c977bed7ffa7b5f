"""The sampling rules of csrc/film_elem.cu (restated in oracle/film_primitives_np.py) against ATen, including the
odd level sizes a 1080p frame produces (135 <- 67, 33 <- 16)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import film as OF
from oracle import film_primitives_np as P


@pytest.mark.parametrize("hw,HW", [((4, 7), (8, 14)), ((67, 30), (135, 60)), ((16, 9), (33, 19)), ((1, 1), (2, 3)),
                                   ((5, 5), (5, 5))])
def test_bilinear_to_size(hw, HW):
    g = torch.Generator().manual_seed(hw[0] * 100 + HW[1])
    v = torch.randn(2, 2, *hw, generator=g)
    ref = F.interpolate(v, size=HW, mode="bilinear")
    # fp32 rounding of the interpolation weights (source coordinates up to ~130 at 24 bits): ~1e-5 on unit-variance data
    assert np.abs(P.bilinear_to_size(v.numpy(), *HW) - ref.numpy()).max() <= 3e-5


@pytest.mark.parametrize("hw,HW", [((4, 7), (8, 14)), ((67, 30), (135, 60)), ((16, 9), (33, 19)), ((8, 8), (17, 16)),
                                   ((5, 5), (5, 5)), ((33, 60), (67, 120))])
def test_nearest_to_size(hw, HW):
    g = torch.Generator().manual_seed(hw[0])
    x = torch.randn(1, 3, *hw, generator=g)
    ref = F.interpolate(x, size=HW, mode="nearest")
    assert np.array_equal(P.nearest_to_size(x.numpy(), *HW), ref.numpy())


@pytest.mark.parametrize("hw", [(8, 6), (135, 7), (3, 67), (2, 2)])
def test_avg_pool2(hw):
    g = torch.Generator().manual_seed(hw[0] + hw[1])
    x = torch.randn(2, 3, *hw, generator=g)
    ref = F.avg_pool2d(x, 2, 2)
    assert np.abs(P.avg_pool2(x.numpy()) - ref.numpy()).max() <= 1e-6


@pytest.mark.parametrize("hw,mag", [((9, 13), 3.0), ((33, 20), 40.0), ((16, 30), 0.4)])
def test_warp_is_pixel_offset_sampling(hw, mag):
    g = torch.Generator().manual_seed(hw[1])
    img = torch.rand(2, 5, *hw, generator=g)
    flow = (torch.rand(2, 2, *hw, generator=g) - 0.5) * 2 * mag   # beyond the border for the large magnitudes
    ref = OF.warp(img, flow)   # the restatement of film_arch.warp that is pinned to the reference's outputs
    got = P.warp_pixel_offsets(img.numpy(), flow.numpy())
    # the reference normalises and un-normalises the coordinates in fp32: ~1e-5 px of coordinate noise
    assert np.abs(got - ref.numpy()).max() <= 2e-4
