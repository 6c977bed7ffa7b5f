"""The numpy restatements of the ATen operators agree with ATen (CPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import primitives_np as P
from oracle import rife46 as O


def _rand(*s, seed=0):
    return np.random.default_rng(seed).standard_normal(s).astype(np.float32)


@pytest.mark.parametrize("s", [0.125, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0])
def test_bilinear_resize(s):
    x = _rand(2, 3, 16, 24)
    ref = F.interpolate(torch.from_numpy(x), scale_factor=s, mode="bilinear", align_corners=False).numpy()
    assert np.abs(P.bilinear_resize(x, s) - ref).max() < 1e-5


def test_downscale_is_two_centre_taps():
    x = _rand(1, 1, 16, 16, seed=3)
    for k in (2, 4, 8):
        ref = F.interpolate(torch.from_numpy(x), scale_factor=1.0 / k, mode="bilinear", align_corners=False).numpy()
        a, b = k // 2 - 1, k // 2
        mine = 0.25 * (x[:, :, a::k, a::k] + x[:, :, a::k, b::k] + x[:, :, b::k, a::k] + x[:, :, b::k, b::k])
        assert np.abs(mine - ref).max() < 1e-6


def test_warp_border():
    img = np.random.default_rng(1).random((2, 3, 20, 28)).astype(np.float32)
    flow = 6 * _rand(2, 2, 20, 28, seed=2)  # plenty of out-of-range coordinates
    ref = O.warp(torch.from_numpy(img), torch.from_numpy(flow)).numpy()
    assert np.abs(P.warp_border(img, flow) - ref).max() < 2e-5  # reference normalise/denormalise rounding


@pytest.mark.parametrize("stride", [1, 2])
def test_conv3x3(stride):
    x, w, b = _rand(2, 5, 12, 10), _rand(7, 5, 3, 3, seed=1), _rand(7, seed=2)
    ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride=stride, padding=1).numpy()
    assert np.abs(P.conv2d_3x3(x, w, b, stride) - ref).max() < 1e-4


def test_conv_transpose_and_shuffle():
    x, w, b = _rand(2, 6, 5, 7), _rand(6, 24, 4, 4, seed=1), _rand(24, seed=2)
    ref = F.conv_transpose2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride=2, padding=1)
    mine = P.conv_transpose2d_k4s2p1(x, w, b)
    assert np.abs(mine - ref.numpy()).max() < 1e-4
    assert np.array_equal(P.pixel_shuffle2(ref.numpy()), F.pixel_shuffle(ref, 2).numpy())
