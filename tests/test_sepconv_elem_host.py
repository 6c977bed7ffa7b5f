"""The element-wise CUDA kernels of csrc/sepconv_elem.cu, compiled for the HOST (tests/host_emu: g++ with a shim that turns
__global__ / blockIdx / __shared__ into plain C++; every grid-stride loop then covers its whole range in one "thread") and
run on the CPU against torch.  This checks the index arithmetic and layouts of the kernels that r01 could not run on a
GPU; it says nothing about launch configurations or the tensor-core kernels."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not (shutil.which("g++") and os.path.exists(os.path.join(CUDA_INC, "cuda_fp16.h"))):
        pytest.skip("g++ / CUDA headers not available")
    so = str(tmp_path_factory.mktemp("emu") / "libsepemu.so")
    src = os.path.join(ROOT, "tests", "host_emu", "sepconv_elem_emu.cpp")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-I" + CUDA_INC, "-o", so, src] + os.environ.get("VFI_EMU_CXXFLAGS", "").split(),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(so)


def h16(t):   # torch float -> contiguous fp16 numpy (NHWC already)
    return np.ascontiguousarray(t.half().numpy())


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def prelu(x, s):
    return torch.where(x > 0, x, s * x)


def test_stats_and_input_conv(emu):
    g = torch.Generator().manual_seed(0)
    H, W, cs = 5, 7, 4                       # odd sizes: replicate pad to 6 x 8; 4-channel frames (alpha ignored)
    He, We = 6, 8
    fr = torch.rand(2, H, W, cs, generator=g)
    frn = np.ascontiguousarray(fr.numpy())
    stats = np.zeros(2, dtype=np.float64)
    emu.emu_stats(p(frn), cs, H, W, He, We, p(stats))
    x = fr[..., :3].permute(0, 3, 1, 2)
    xp = F.pad(x, [0, We - W, 0, He - H], mode="replicate")
    assert abs(stats[0] - float(xp.double().sum())) < 1e-9 * xp.numel() + 1e-6
    assert abs(stats[1] - float((xp.double() ** 2).sum())) < 1e-6
    w = (torch.rand(16, 3, 3, 3, generator=g) - 0.5)
    b = (torch.rand(16, generator=g) - 0.5)
    slope = 0.3
    out = np.zeros((He // 2) * (We // 2) * 128, dtype=np.float16)
    emu.emu_input_conv(p(frn), cs, H, W, He, We, p(stats), p(np.ascontiguousarray(w.numpy())), p(np.ascontiguousarray(b.numpy())),
                       C.c_float(slope), p(out))
    # reference: sepconv_enhanced.py:620-642 + the PReLU of netEncode.0.netVer.1, then space-to-depth with channel = frame*16+c
    mean = xp.reshape(1, -1).mean(1).view(1, 1, 1, 1)
    std = xp.reshape(1, -1).std(1).view(1, 1, 1, 1)
    feat = torch.cat([F.conv2d((xp[k:k + 1] - mean) / (std + 1e-7), w, b, padding=1) for k in (0, 1)], 1)
    feat = prelu(feat, slope)[0]             # [32, He, We]
    exp = torch.stack([feat[:, a::2, bb::2] for a in (0, 1) for bb in (0, 1)], 0)   # [4, 32, He/2, We/2]
    exp = exp.permute(2, 3, 0, 1).reshape(-1)
    assert np.abs(out.astype(np.float32) - exp.numpy()).max() <= 4e-3


@pytest.mark.parametrize("hw", [(6, 8), (5, 7), (1, 3)])
def test_prelu_s2d16(emu, hw):
    g = torch.Generator().manual_seed(hw[0])
    B, Cc = 2, 16
    x = torch.randn(B, *hw, Cc, generator=g)
    xin = h16(x)
    ho, wo = (hw[0] + 1) // 2, (hw[1] + 1) // 2
    out = np.full(B * ho * wo * 4 * Cc, 7, dtype=np.float16)
    emu.emu_prelu_s2d16(p(xin), p(out), C.c_float(0.2), Cc, B, hw[0], hw[1])
    xr = prelu(torch.from_numpy(xin).float(), 0.2)
    xp = F.pad(xr.permute(0, 3, 1, 2), (0, 2 * wo - hw[1], 0, 2 * ho - hw[0]))
    exp = torch.stack([xp[:, :, a::2, bb::2] for a in (0, 1) for bb in (0, 1)], 1)   # [B, 4, C, ho, wo]
    exp = exp.permute(0, 3, 4, 1, 2).reshape(-1)
    assert np.array_equal(out.astype(np.float32), exp.half().float().numpy())


def test_prelu16_and_add_crop(emu):
    g = torch.Generator().manual_seed(3)
    x = h16(torch.randn(2, 3, 5, 8, generator=g))
    out = np.zeros_like(x)
    emu.emu_prelu16(p(x), p(out), C.c_float(-0.5), C.c_size_t(x.size))
    assert np.array_equal(out, prelu(torch.from_numpy(x).float(), -0.5).half().numpy())
    v = h16(torch.randn(2, 4, 6, 8, generator=g))
    acc = x.copy()
    emu.emu_add_crop16(p(v), 4, 6, p(acc), 8, 2, 3, 5)
    exp = (torch.from_numpy(x).float() + torch.from_numpy(v).float()[:, :3, :5]).half().numpy()
    assert np.array_equal(acc, exp)


@pytest.mark.parametrize("hw,tgt", [((3, 4), (6, 8)), ((3, 4), (5, 7)), ((1, 1), (2, 2))])
def test_prelu_up2_16(emu, hw, tgt):
    g = torch.Generator().manual_seed(tgt[0])
    B, Cc = 2, 8
    x = h16(torch.randn(B, *hw, Cc, generator=g))
    out = np.zeros(B * tgt[0] * tgt[1] * Cc, dtype=np.float16)
    emu.emu_prelu_up2_16(p(x), p(out), C.c_float(0.25), Cc, B, hw[0], hw[1], tgt[0], tgt[1])
    ref = F.interpolate(prelu(torch.from_numpy(x).float(), 0.25).permute(0, 3, 1, 2), scale_factor=2.0, mode="bilinear",
                        align_corners=False)[:, :, :tgt[0], :tgt[1]].permute(0, 2, 3, 1).reshape(-1)
    assert np.abs(out.astype(np.float32) - ref.numpy()).max() <= 3e-3


def test_coeff_pad_finish(emu):
    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 4, 6
    k = h16(torch.randn(B, H, W, 64, generator=g))
    out = np.zeros(B * 51 * H * W, dtype=np.float32)
    emu.emu_coeff_nchw(p(k), 64, p(out), 51, B, H, W)
    exp = torch.from_numpy(k).float()[..., :51].permute(0, 3, 1, 2).reshape(-1)
    assert np.array_equal(out, exp.numpy())
    # frame -> replicate pad to even, then by 25, + ones channel (sepconv_enhanced.py:611-618, :644-681)
    h0, w0, cs = 5, 7, 3
    fr = torch.rand(2, h0, w0, cs, generator=g)
    frn = np.ascontiguousarray(fr.numpy())
    He, We = 6, 8
    Hp, Wp = He + 50, We + 50
    for which in (0, 1):
        o = np.zeros(4 * Hp * Wp, dtype=np.float32)
        emu.emu_pad_input(p(frn), cs, which, h0, w0, Hp, Wp, p(o))
        x = fr[which:which + 1].permute(0, 3, 1, 2)
        xp = F.pad(F.pad(x, [0, We - w0, 0, He - h0], mode="replicate"), [25] * 4, mode="replicate")
        exp = torch.cat([xp, torch.ones(1, 1, Hp, Wp)], 1).reshape(-1)
        assert np.array_equal(o, exp.numpy())
    o1 = torch.rand(B, 4, He, We, generator=g)
    o2 = torch.rand(B, 4, He, We, generator=g)
    o1[0, 3, 0, 0], o2[0, 3, 0, 0] = 0.004, 0.001   # |normaliser| < 0.01 -> 1 (:697-698)
    res = np.zeros(B * h0 * w0 * 3, dtype=np.float32)
    emu.emu_finish(p(np.ascontiguousarray(o1.numpy())), p(np.ascontiguousarray(o2.numpy())), p(res), B, h0, w0, He, We)
    s = o1 + o2
    n = s[:, 3:4].clone()
    n[n.abs() < 0.01] = 1.0
    exp = (s[:, :3] / n)[:, :, :h0, :w0].permute(0, 2, 3, 1).reshape(-1)
    assert np.abs(res - exp.numpy()).max() <= 1e-6
