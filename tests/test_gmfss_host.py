"""The GMFSS Fortuna path on the CPU: csrc/gmops.cu compiled for the host (tests/host_emu/gmops_emu.cpp, same vfi_gm_* C ABI)
under the product's own schedule (comfyui-frame-interpolation_b200/gmfss.py) against outputs of the UNMODIFIED reference model
(tests/golden/gmfss_*.npz).  The two components with their own GPU tests - the fused soft splat and the RIFE engine - are the
oracle's here; everything else (GMFlow with its encoder / swin transformer / matching / propagation / convex up-sampling,
MetricNet, FeatureNet, GridNet, the resampling and the glue) is the product's kernels."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_gmfss import gmfss_cases  # noqa: E402
from oracle import film as OF  # noqa: E402
from oracle import gmflow as GF  # noqa: E402
from oracle import gmfss as OG  # noqa: E402
from oracle import gmfss_weights as GW  # noqa: E402
from oracle import ops_ref  # noqa: E402
from oracle import rife46 as R  # noqa: E402

CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def model_factory(tmp_path_factory):
    if not (shutil.which("g++") and os.path.exists(os.path.join(CUDA_INC, "cuda_fp16.h"))):
        pytest.skip("g++ / CUDA headers not available")
    so = str(tmp_path_factory.mktemp("emu") / "libgmopsemu.so")
    src = os.path.join(ROOT, "tests", "host_emu", "gmops_emu.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++20", "-shared", "-fPIC", "-pthread", "-I" + CUDA_INC, "-o", so, src] +
                       os.environ.get("VFI_EMU_CXXFLAGS", "").split(), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    import __graft_entry__ as ge
    ge.load_package()
    from cfi_b200.gmfss import GMFSS, Ops

    def make(sds):
        o = Ops(lib, torch.device("cpu"))
        splat = lambda x, fl, z: ops_ref.softsplat(x, fl, z, "soft")   # noqa: E731
        rife = lambda h0, h1, t: R.ifnet46_forward(sds["ifnet"], h0, h1, torch.full((h0.shape[0], 1, 1, 1), float(t)), (8, 4, 2, 1))  # noqa: E731
        return GMFSS(sds, o, splat=splat, rife=rife)
    return make


def _case(name):
    cfg = gmfss_cases()[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sds = GW.synthetic_state_dicts(cfg["seed"])
    fr = OF.synthetic_clip(2, cfg["h"], cfg["w"], seed=cfg["clip_seed"]).permute(0, 3, 1, 2).contiguous()
    return cfg, g, sds, fr


def test_primitives_against_torch(model_factory):
    sds = GW.synthetic_state_dicts(0)
    o = model_factory(sds).o
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 13, 18, generator=g)
    w = torch.randn(7, 5, 3, 3, generator=g) * 0.2
    b = torch.randn(7, generator=g)
    r1 = torch.randn(2, 7, 7, 9, generator=g)
    ref = F.conv2d(F.prelu(x, torch.tensor([0.3])), w, b, stride=2, padding=1) + r1
    assert (o.conv(x, w, b, stride=2, pre=0.3, res1=r1) - ref).abs().max() <= 1e-5
    w7 = torch.randn(4, 5, 7, 7, generator=g) * 0.1
    assert (o.conv(x, w7, stride=2, pad=3) - F.conv2d(x, w7, None, stride=2, padding=3)).abs().max() <= 1e-5
    # the packed-weight fast forms (3x3 s1 / s2, 1x1 s2, 7x7 s2; Cout not a multiple of 16; channel slices in and out)
    for (kk, st, pd) in ((3, 1, 1), (3, 2, 1), (1, 2, 0), (1, 1, 0), (7, 2, 3)):
        wk = torch.randn(19, 5, kk, kk, generator=g) * 0.2
        o.register_conv(wk)
        big_in = torch.randn(2, 9, 13, 18, generator=g)
        ref = F.conv2d(F.prelu(big_in[:, 3:8], torch.tensor([0.3])), wk, b.repeat(3)[:19], stride=st, padding=pd)
        dst = torch.zeros(2, 25, ref.shape[2], ref.shape[3])
        rr = torch.randn_like(ref)
        o.conv(big_in, wk, b.repeat(3)[:19].contiguous(), stride=st, pad=pd, pre=0.3, res1=rr, out=dst, out_coff=4, in_coff=3, cin=5, post=1)
        assert (dst[:, 4:23] - F.relu(ref + rr)).abs().max() <= 1e-5, (kk, st)
        assert float(dst[:, :4].abs().max()) == 0.0 and float(dst[:, 23:].abs().max()) == 0.0
    wl8 = torch.randn(24, 16, generator=g)
    o.register_linear(wl8)
    t21 = torch.randn(3, 7, 16, generator=g)       # 21 rows: a partial 8-row tile
    assert (o.linear(t21, wl8, torch.arange(24.0), act=1) - F.gelu(F.linear(t21, wl8, torch.arange(24.0)))).abs().max() <= 1e-5
    for rows in (60001, 40003, 37):            # 8, 4 and 2 rows per thread in the tiled product (a partial last tile each)
        aa, bb8 = torch.randn(1, rows, 8, generator=g), torch.randn(1, 8, 64, generator=g)
        assert (o.bmm(aa, bb8, False, 0.25) - 0.25 * aa @ bb8).abs().max() <= 1e-5, rows
    qq, kk2 = torch.randn(4, 10, 16, generator=g), torch.randn(4, 24, 16, generator=g)
    mm = torch.randn(2, 10, 24, generator=g)
    assert (o.scores(qq, kk2, 0.5, mm) - (0.5 * qq @ kk2.transpose(1, 2) + mm.repeat(2, 1, 1))).abs().max() <= 1e-5
    wt = torch.randn(5, 6, 4, 4, generator=g) * 0.2
    bt = torch.randn(6, generator=g)
    ref = F.conv_transpose2d(F.prelu(x, torch.tensor([0.2])), wt, bt, stride=2, padding=1)
    assert (o.convt4(x, wt, bt, pre=0.2) - ref).abs().max() <= 1e-5
    assert (o.inorm(x, relu=True) - F.relu(F.instance_norm(x, eps=1e-5))).abs().max() <= 1e-5
    t = torch.randn(3, 40, 16, generator=g)
    gm, bs = torch.randn(16, generator=g), torch.randn(16, generator=g)
    assert (o.layer_norm(t, gm, bs, src=t) - (t + F.layer_norm(t, (16,), gm, bs))).abs().max() <= 1e-5
    wl = torch.randn(24, 16, generator=g)
    assert (o.linear(t, wl, act=1) - F.gelu(F.linear(t, wl))).abs().max() <= 1e-5
    a, bb = torch.randn(4, 10, 16, generator=g), torch.randn(4, 12, 16, generator=g)
    m = torch.randn(2, 10, 12, generator=g)
    assert (o.bmm(a, bb, True, 0.5, m) - (0.5 * a @ bb.transpose(1, 2) + m.repeat(2, 1, 1))).abs().max() <= 1e-5
    s = torch.randn(6, 33, generator=g)
    assert (o.softmax_(s.clone()) - torch.softmax(s, -1)).abs().max() <= 1e-6
    fl = torch.randn(2, 2, 13, 18, generator=g) * 3
    assert (o.warp(x, fl) - OG._flow_warp(x, fl)).abs().max() <= 1e-5
    assert (o.resize(x, 6, 9) - F.interpolate(x, size=(6, 9), mode="bilinear", align_corners=False)).abs().max() <= 1e-5
    assert (o.resize(x, 26, 36, align=True, mul=2.0) - 2 * F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)).abs().max() <= 1e-5
    ps = torch.randn(2, 12, 5, 6, generator=g)
    assert torch.equal(o.pixel_shuffle2(ps), F.pixel_shuffle(ps, 2))
    # window split / merge with the cyclic shift, position embedding, matching, propagation, convex up-sampling
    f0, f1 = torch.randn(1, 16, 8, 12, generator=g), torch.randn(1, 16, 8, 12, generator=g)
    a0, a1 = GF.add_position(f0, f1, 2)
    assert (o.add_position_(f0.clone(), 2) - a0).abs().max() <= 1e-5
    tok = o.to_tokens(f0)
    assert torch.equal(tok, f0.flatten(-2).permute(0, 2, 1).contiguous())
    rolled = torch.roll(tok.view(1, 8, 12, 16), shifts=(-2, -3), dims=(1, 2))
    assert torch.equal(o.window(tok, 2, 2, 3, True, 1, 8, 12, 16), GF._windows_last(rolled, 2).reshape(4, -1, 16))
    assert torch.equal(o.window(o.window(tok, 2, 2, 3, True, 1, 8, 12, 16), 2, 2, 3, False, 1, 8, 12, 16), tok)
    assert (o.local_match(f0, f1, 4) - GF.local_match(f0, f1, 4)).abs().max() <= 1e-4
    mask = torch.randn(1, 9 * 16, 8, 12, generator=g)
    flow = torch.randn(1, 2, 8, 12, generator=g)
    b_, _, h_, w_ = flow.shape
    mk = torch.softmax(mask.view(b_, 1, 9, 4, 4, h_, w_), dim=2)
    up = F.unfold(4 * flow, [3, 3], padding=1).view(b_, 2, 9, 1, 1, h_, w_)
    ref = torch.sum(mk * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(b_, 2, 4 * h_, 4 * w_)
    assert (o.convex_up(mask, flow, 4) - ref).abs().max() <= 1e-5
    i0, i1 = torch.rand(1, 3, 8, 12, generator=g), torch.rand(1, 3, 8, 12, generator=g)
    fa, fb = torch.randn(1, 2, 8, 12, generator=g) * 2, torch.randn(1, 2, 8, 12, generator=g) * 2
    m0 = F.l1_loss(i0, OG.backwarp(i1, fa), reduction="none").mean([1], True)
    occ_f, occ_b = OG.fb_consistency(fa, fb)
    mi = o.metric_input(i0, i1, fa, fb)
    assert (mi[:, 6:7] + m0).abs().max() <= 1e-5 and torch.equal(mi[:, 12], occ_f) and torch.equal(mi[:, 13], occ_b)


@pytest.mark.parametrize("name", sorted(gmfss_cases().keys()))
def test_gmflow_on_the_product_kernels(model_factory, name):
    cfg, g, sds, fr = _case(name)
    h, w = cfg["h"], cfg["w"]
    ph, pw = ((h - 1) // 64 + 1) * 64, ((w - 1) // 64 + 1) * 64
    i0, i1 = F.pad(fr[0:1], (0, pw - w, 0, ph - h)), F.pad(fr[1:2], (0, pw - w, 0, ph - h))
    h0 = F.interpolate(i0, scale_factor=0.5, mode="bilinear", align_corners=False).contiguous()
    h1 = F.interpolate(i1, scale_factor=0.5, mode="bilinear", align_corners=False).contiguous()
    m = model_factory(sds)
    f01 = m.gmflow(h0, h1)
    assert f01.shape == g["flow01"].shape
    err = (f01 - torch.from_numpy(g["flow01"])).abs().max().item()
    print(f"{name}: gmflow max abs error {err:.2e} px")
    assert err <= 5e-3


@pytest.mark.parametrize("name", sorted(gmfss_cases().keys())[:1])
def test_whole_model_on_the_product_kernels(model_factory, name):
    """frames -> frame through GMFSS.interpolate (padding, reuse, inference, crop) vs the unmodified reference's frame"""
    cfg, g, sds, fr = _case(name)
    m = model_factory(sds)
    out = m.interpolate(fr[0:1].contiguous(), fr[1:2].contiguous(), cfg["t"])
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape
    psnr = R.psnr(out, ref)
    print(f"{name}: whole GMFSS on the product kernels (host build), PSNR {psnr:.1f} dB, max abs {(out - ref).abs().max().item():.2e}")
    assert psnr >= 60.0
