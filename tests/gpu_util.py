"""Helpers shared by the -m gpu tests (layout conversions between the oracle's NCHW fp32 and the kernels' layouts)."""
import torch


def s2d(x):
    """[B,H,W,C] -> [B,H/2,W/2,4C], channel order (row parity a, column parity b, c) - the layout conv0.x reads."""
    b, h, w, c = x.shape
    return x.view(b, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h // 2, w // 2, 4 * c).contiguous()


def un_s2d(y, c):
    b, h2, w2, _ = y.shape
    return y.view(b, h2, w2, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h2 * 2, w2 * 2, c).contiguous()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
