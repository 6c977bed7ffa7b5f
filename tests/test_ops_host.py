"""csrc/ops.cu (the kernels behind vfi_softsplat_sum / vfi_costvol_l1 / vfi_corr_dot / vfi_sepconv, incl. the shared-memory
tiled ones) compiled for the HOST (tests/host_emu/cuda_shim_block.h: blocks on host threads, threads as fibers) against the
oracle restatements that tests/test_ops_ref_pinned.py pins to the reference's own kernels.  The same kernels are GPU-tested
in tests/test_gpu_ops.py; this keeps them under test where there is no GPU."""
import ctypes as C
import os
import shutil
import subprocess

import pytest
import torch

from oracle import ops_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not (shutil.which("g++") and os.path.exists(os.path.join(CUDA_INC, "cuda_fp16.h"))):
        pytest.skip("g++ / CUDA headers not available")
    so = str(tmp_path_factory.mktemp("emu") / "libopsemu.so")
    src = os.path.join(ROOT, "tests", "host_emu", "ops_emu.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++20", "-shared", "-fPIC", "-pthread", "-I" + CUDA_INC, "-o", so, src] + os.environ.get("VFI_EMU_CXXFLAGS", "").split(),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(so)


def vp(t):
    return C.c_void_p(t.data_ptr())


def test_softsplat_kernel(emu):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 19, 23, generator=g)
    flow = (torch.rand(2, 2, 19, 23, generator=g) - 0.5) * 9
    out = torch.full_like(x, 7.0)   # the launcher zeroes the output itself
    assert emu.emu_softsplat(vp(x), vp(flow), vp(out), 2, 5, 19, 23) == 0
    assert (out - ops_ref.softsplat_sum(x, flow)).abs().max().item() <= 1e-4   # atomics: summation order differs


@pytest.mark.parametrize("dot", [0, 1])
def test_volume81_kernels(emu, dot):
    g = torch.Generator().manual_seed(dot)
    a, b = torch.randn(1, 20, 21, 37, generator=g), torch.randn(1, 20, 21, 37, generator=g)
    out = torch.zeros(1, 81, 21, 37)
    assert emu.emu_volume81(dot, vp(a), vp(b), vp(out), 1, 20, 21, 37) == 0
    ref = ops_ref.correlation_dot(a, b) if dot else ops_ref.costvol_l1(a, b)
    assert (out - ref).abs().max().item() <= 1e-4


def test_sepconv_tiled_kernel(emu):
    g = torch.Generator().manual_seed(3)
    h, w, k = 19, 45, 51                       # not multiples of the 16 x 32 tile
    x = torch.randn(1, 4, h + k - 1, w + k - 1, generator=g)
    ver, hor = torch.randn(1, k, h, w, generator=g), torch.randn(1, k, h, w, generator=g)
    out = torch.zeros(1, 4, h, w)
    assert emu.emu_sepconv(vp(x), vp(ver), vp(hor), vp(out), 1, 4, h, w, k, k) == 0
    ref = ops_ref.sepconv(x, ver, hor)
    assert (out - ref).abs().max().item() <= 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("mode", ["avg", "linear", "soft", "soft-zeroeps", "linear-clipeps"])
def test_softsplat_fused_modes(emu, mode):
    """vfi_softsplat_weighted (two launches, no temporaries) == the reference's cat -> splat -> divide chain."""
    g = torch.Generator().manual_seed(len(mode))
    x = torch.randn(2, 6, 17, 21, generator=g)
    flow = (torch.rand(2, 2, 17, 21, generator=g) - 0.5) * 8
    metric = torch.rand(2, 1, 17, 21, generator=g) * 2 - (0.5 if mode.startswith("soft") else -0.2)
    base = mode.split("-")[0]
    m = {"avg": 0, "linear": 1, "soft": 2}[base]
    e = {"addeps": 0, "zeroeps": 1, "clipeps": 2}[mode.split("-")[1] if "-" in mode else "addeps"]
    out, norm = torch.full_like(x, 3.0), torch.full((2, 1, 17, 21), 3.0)
    assert emu.emu_softsplat_weighted(vp(x), vp(flow), None if m == 0 else vp(metric), m, e, vp(out), vp(norm), 2, 6, 17, 21) == 0
    ref = ops_ref.softsplat(x, flow, None if m == 0 else metric, mode)
    ok = ref.abs() < 1e4                       # addeps: untouched targets are 0 / 1e-7 in both, tiny norms blow up alike
    assert (out - ref)[ok].abs().max().item() <= 2e-3 * max(1.0, float(ref[ok].abs().max()))


@pytest.mark.parametrize("mode", ["avg", "linear", "soft", "soft-zeroeps", "linear-clipeps"])
@pytest.mark.parametrize("c", [6, 3])
def test_softsplat_channel_last_modes(emu, mode, c):
    """the channel-last form behind vfi_softsplat_weighted since r02 (NCHW -> NHWC transpose, 16-byte vector atomics per four
    channels, normalise + transpose back) == the reference's chain; C = 3 exercises the channel padding"""
    g = torch.Generator().manual_seed(len(mode) + c)
    x = torch.randn(2, c, 17, 21, generator=g)
    flow = (torch.rand(2, 2, 17, 21, generator=g) - 0.5) * 8
    metric = torch.rand(2, 1, 17, 21, generator=g) * 2 - (0.5 if mode.startswith("soft") else -0.2)
    base = mode.split("-")[0]
    m = {"avg": 0, "linear": 1, "soft": 2}[base]
    e = {"addeps": 0, "zeroeps": 1, "clipeps": 2}[mode.split("-")[1] if "-" in mode else "addeps"]
    out, norm = torch.full_like(x, 3.0), torch.full((2, 1, 17, 21), 3.0)
    cp = (c + 3) // 4 * 4
    sa, sb = torch.full((2 * 17 * 21 * cp,), 5.0), torch.full((2 * 17 * 21 * cp,), 5.0)
    assert emu.emu_softsplat_weighted_nhwc(vp(x), vp(flow), None if m == 0 else vp(metric), m, e, vp(out), vp(norm), vp(sa), vp(sb), 2, c,
                                           17, 21) == 0
    ref = ops_ref.softsplat(x, flow, None if m == 0 else metric, mode)
    ok = ref.abs() < 1e4
    assert (out - ref)[ok].abs().max().item() <= 2e-3 * max(1.0, float(ref[ok].abs().max()))


@pytest.mark.parametrize("dot", [0, 1])
def test_volume81_warp_kernels(emu, dot):
    """the warp-per-pixel channel-last volume (shuffle reduction over the channels) == the oracle; C = 20 is not a multiple of
    32 (partial lane coverage) nor of 4 (channel padding), W = 37 not a multiple of the 8 pixels of a block"""
    g = torch.Generator().manual_seed(10 + dot)
    a, b = torch.randn(1, 20, 13, 37, generator=g), torch.randn(1, 20, 13, 37, generator=g)
    out = torch.zeros(1, 81, 13, 37)
    sa, sb = torch.zeros(13 * 37 * 20), torch.zeros(13 * 37 * 20)
    assert emu.emu_volume81_warp(dot, vp(a), vp(b), vp(out), vp(sa), vp(sb), 1, 20, 13, 37) == 0
    ref = ops_ref.correlation_dot(a, b) if dot else ops_ref.costvol_l1(a, b)
    assert (out - ref).abs().max().item() <= 1e-4
