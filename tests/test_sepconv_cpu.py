"""Sepconv host side on CPU: the stride-2 conv of the encoder as streamconv runs it (space-to-depth input, 2x2 taps with
the padding row / column before) decoded from the library's own packer and checked against torch's stride-2 conv, odd
sizes included; the up-sampling rule of sepconv_elem.cu against ATen; node surface; loud error without a GPU."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sepconv as OS


def _decode(packed, ksize, ktot, n_total, n_cta, nsplit):
    ntaps = ksize * ksize
    nkb = ktot // 64
    per_split = nkb * ntaps * n_cta * 64
    p = packed[:nsplit * per_split].view(np.float16).astype(np.float32).reshape(nsplit, nkb, ntaps, n_cta, 8, 8)
    w = np.zeros((n_total, ktot, ntaps), dtype=np.float32)
    for nl in range(n_cta):
        for chunk in range(8):
            src = p[:, :, :, nl, chunk ^ (nl & 7), :]
            for sp in range(nsplit):
                for kb in range(nkb):
                    w[sp * n_cta + nl, kb * 64 + chunk * 8: kb * 64 + chunk * 8 + 8, :] = src[sp, kb].T
    return torch.from_numpy(w.reshape(n_total, ktot, ksize, ksize))


def _s2d(x):
    """prelu_s2d16_kernel's layout (slope 1): [B,C,H,W] -> [B, (a, b, C), ceil(H/2), ceil(W/2)], zero past an odd edge."""
    b, c, h, w = x.shape
    ho, wo = (h + 1) // 2, (w + 1) // 2
    xp = F.pad(x, (0, 2 * wo - w, 0, 2 * ho - h))
    return torch.cat([xp[:, :, a::2, bb::2] for a in (0, 1) for bb in (0, 1)], 1)


@pytest.mark.parametrize("cin,cout,hw", [(32, 64, (12, 16)), (64, 128, (9, 7)), (128, 256, (5, 6))])
def test_stride2_conv_as_s2d_2x2(pkg, cin, cout, hw):
    from cfi_b200._lib import lib
    L = lib()
    g = torch.Generator().manual_seed(cin + hw[0])
    w = (torch.rand(cout, cin, 3, 3, generator=g) - 0.5).half().float()
    wn = np.ascontiguousarray(w.numpy())
    cap = 16 * 1024 * 1024
    out = np.zeros(cap, dtype=np.uint16)
    c0, n_cta, nsplit = C.c_int(), C.c_int(), C.c_int()
    rc = L.vfi_sepconv_debug_pack_host(1, cin, cout, cout, 0, wn.ctypes.data, out.ctypes.data, cap, C.byref(c0),
                                       C.byref(n_cta), C.byref(nsplit))
    assert rc == 0, L.vfi_last_error()
    assert c0.value == 4 * cin
    wd = _decode(out, 2, c0.value, cout, n_cta.value, nsplit.value)       # [cout, 4 cin, 2, 2]
    x = torch.rand(2, cin, *hw, generator=g) - 0.5
    ref = F.conv2d(x, w, stride=2, padding=1)
    # streamconv k = 2 with pad_before = 1: out[y, x] = sum_{ty, tx} W[ty, tx] . in[y + ty - 1, x + tx - 1], zero outside
    got = F.conv2d(F.pad(_s2d(x), (1, 0, 1, 0)), wd)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-4


def test_stride1_pack_matches_film_packer(pkg):
    from cfi_b200._lib import lib
    L = lib()
    g = torch.Generator().manual_seed(9)
    w = (torch.rand(51, 64, 3, 3, generator=g) - 0.5).half().float()
    wn = np.ascontiguousarray(w.numpy())
    cap = 1 << 22
    a, b = np.zeros(cap, dtype=np.uint16), np.zeros(cap, dtype=np.uint16)
    i = [C.c_int() for _ in range(4)]
    assert L.vfi_sepconv_debug_pack_host(0, 64, 51, 64, 0, wn.ctypes.data, a.ctypes.data, cap, C.byref(i[0]), C.byref(i[1]),
                                         C.byref(i[2])) == 0
    assert L.vfi_film_debug_pack_host(0, 0, 0, 3, 64, 0, wn.ctypes.data, 51, 64, b.ctypes.data, cap, C.byref(i[0]),
                                      C.byref(i[3]), C.byref(i[1]), C.byref(i[2])) == 0
    assert np.array_equal(a, b)


@pytest.mark.parametrize("hw,crop", [((6, 8), (12, 16)), ((68, 5), (135, 9)), ((3, 4), (5, 8))])
def test_up2_rule(hw, crop):
    """prelu_up2_16_kernel: src = max((dst + 0.5) * 0.5 - 0.5, 0), second tap clamped, then crop."""
    g = torch.Generator().manual_seed(hw[0])
    x = torch.randn(1, 3, *hw, generator=g)
    ref = F.interpolate(x, scale_factor=2.0, mode="bilinear", align_corners=False)[:, :, :crop[0], :crop[1]]

    def taps(o, i):
        src = np.maximum((np.arange(o, dtype=np.float32) + np.float32(0.5)) * np.float32(0.5) - np.float32(0.5), 0)
        i0 = np.minimum(src.astype(np.int64), i - 1)
        return i0, np.minimum(i0 + 1, i - 1), (src - i0).astype(np.float32)

    y0, y1, ly = taps(crop[0], hw[0])
    x0, x1, lx = taps(crop[1], hw[1])
    v = x.numpy()
    top = v[:, :, y0][:, :, :, x0] * (1 - lx) + v[:, :, y0][:, :, :, x1] * lx
    bot = v[:, :, y1][:, :, :, x0] * (1 - lx) + v[:, :, y1][:, :, :, x1] * lx
    got = top * (1 - ly)[None, None, :, None] + bot * ly[None, None, :, None]
    assert np.abs(got - ref.numpy()).max() <= 2e-6


def test_sepconv_names_and_node_surface(pkg):
    import cfi_b200 as P
    import cfi_b200.sepconv_node as SN
    from cfi_b200.engine import sepconv_state_dict_names
    assert sepconv_state_dict_names() == [n for n, _ in OS.state_dict_spec()]
    assert P.NODE_CLASS_MAPPINGS["Sepconv VFI"] is SN.SepconvVFI
    it = SN.SepconvVFI.INPUT_TYPES()   # sepconv/__init__.py:15-30
    assert list(it["required"].keys()) == ["ckpt_name", "frames", "clear_cache_after_n_frames", "multiplier"]
    assert it["required"]["ckpt_name"] == (["sepconv.pth"],)
    assert it["required"]["multiplier"][1] == {"default": 2, "min": 2, "max": 1000}
    assert SN.SepconvVFI.RETURN_TYPES == ("IMAGE",) and SN.SepconvVFI.FUNCTION == "vfi"


def test_sepconv_node_runs_the_bisection_loop(pkg):
    """The node through generic_frame_loop(use_timestep=False) with a stand-in engine (tests only): multiplier 4 = three
    recursive midpoints per pair in output order, originals passed through."""
    import cfi_b200.sepconv_node as SN

    class Mid:
        def middle_frame(self, a, b):
            return 0.5 * (a + b)

    fr = torch.rand(3, 6, 8, 4, generator=torch.Generator().manual_seed(1))
    (out,) = SN.SepconvVFI().vfi("sepconv.pth", fr, multiplier=4, _engine=Mid())
    assert out.shape == (9, 6, 8, 3)
    x = fr[..., :3]
    for p in range(2):
        for k in range(4):
            exp = (1 - k / 4) * x[p] + (k / 4) * x[p + 1]
            assert (out[4 * p + k] - exp).abs().max().item() <= 1e-6
    assert torch.equal(out[8], x[2])


def test_sepconv_no_gpu_is_a_loud_error(pkg):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cfi_b200._lib import VfiError
    from cfi_b200.engine import SepconvEngine
    with pytest.raises(VfiError):
        SepconvEngine(OS.synthetic_state_dict(0), 0)
