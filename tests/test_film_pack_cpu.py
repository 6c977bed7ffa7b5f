"""Host-side weight packing of the FILM path (csrc/film.cu pack_conv + channel maps) on CPU: the packed tensor-core
operand is decoded exactly as csrc/streamconv.cu reads it (streamconv_ref_kernel indexing: [split][k-block][tap][row]
[64 channels, 16-byte chunks XOR (row & 7)], tap = ky * k + kx at offset (ky - (k-1)//2, kx - (k-1)//2)) and the
convolution evaluated from the decoded weights on our padded channel layout must equal torch's conv2d(padding='same')
on the reference's channel order.  No GPU, no CUDA calls."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _pack(L, layout, Cf, nf, ksize, n_total, w):
    cout, cin = w.shape[0], w.shape[1]
    w = np.ascontiguousarray(w.numpy().astype(np.float32))
    cap = 64 * 1024 * 1024
    out = np.zeros(cap, dtype=np.uint16)
    c0, c1, n_cta, nsplit = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = L.vfi_film_debug_pack_host(layout, Cf, nf, ksize, n_total, 0, w.ctypes.data, cout, cin, out.ctypes.data, cap,
                                    C.byref(c0), C.byref(c1), C.byref(n_cta), C.byref(nsplit))
    assert rc == 0, L.vfi_last_error()
    return out, c0.value, c1.value, n_cta.value, nsplit.value


def _decode(packed, ksize, ktot, n_total, n_cta, nsplit):
    """-> fp32 weights [n_total, ktot (padded input channels), k, k] as the kernels see them."""
    ntaps = ksize * ksize
    nkb = ktot // 64
    per_split = nkb * ntaps * n_cta * 64
    p = packed[:nsplit * per_split].view(np.float16).astype(np.float32).reshape(nsplit, nkb, ntaps, n_cta, 8, 8)
    w = np.zeros((n_total, ktot, ntaps), dtype=np.float32)
    for nl in range(n_cta):
        for chunk in range(8):
            src = p[:, :, :, nl, chunk ^ (nl & 7), :]          # [split, kb, tap, 8]
            for sp in range(nsplit):
                for kb in range(nkb):
                    w[sp * n_cta + nl, kb * 64 + chunk * 8: kb * 64 + chunk * 8 + 8, :] = src[sp, kb].T
    return torch.from_numpy(w.reshape(n_total, ktot, ksize, ksize))


def _conv_same(x, w):
    """out[y, x] = sum w[ky, kx] * in[y + ky - (k-1)//2, x + kx - (k-1)//2], zero outside: the kernels' tap geometry."""
    k = w.shape[-1]
    lo = (k - 1) // 2
    hi = k - 1 - lo
    return F.conv2d(F.pad(x, (lo, hi, lo, hi)), w)


@pytest.mark.parametrize("ksize,cin,cout,n_total", [(3, 128, 128, 128), (3, 32, 32, 64), (1, 32, 16, 16), (2, 128, 64, 64),
                                                    (3, 256, 512, 512)])
def test_pack_identity_layout(pkg, ksize, cin, cout, n_total):
    from cfi_b200._lib import lib
    L = lib()
    g = torch.Generator().manual_seed(ksize * 1000 + cin)
    w = (torch.rand(cout, cin, ksize, ksize, generator=g) - 0.5).half().float()
    packed, c0, c1, n_cta, nsplit = _pack(L, 0, 0, 0, ksize, n_total, w)
    assert c0 == (cin + 63) // 64 * 64 and c1 == 0 and n_cta * nsplit == n_total and n_cta <= 128
    wd = _decode(packed, ksize, c0, n_total, n_cta, nsplit)
    assert torch.equal(wd[:cout, :cin], w)
    assert float(wd[cout:].abs().max()) == 0 if n_total > cout else True
    assert float(wd[:, cin:].abs().max()) == 0 if c0 > cin else True
    x = torch.rand(1, cin, 9, 7, generator=g) - 0.5
    xp = torch.cat([x, torch.full((1, c0 - cin, 9, 7), 3.0)], 1) if c0 > cin else x   # padding channels are multiplied by 0
    ref = F.conv2d(x, w, padding="same")
    assert (_conv_same(xp, wd)[:, :cout] - ref).abs().max().item() <= 1e-4


@pytest.mark.parametrize("Cf,nf,ksize", [(64, 64, 3), (192, 128, 3), (64, 0, 2)])
def test_pack_fusion_layout(pkg, Cf, nf, ksize):
    """[wfeat0 C | wfeat1 C | misc 64] (+ decoder state) vs the reference's [wimg0, wfeat0, wimg1, wfeat1, bwd, fwd] (+ net)
    (film_arch.py:431-447, :292)."""
    from cfi_b200._lib import lib
    L = lib()
    g = torch.Generator().manual_seed(Cf + nf + ksize)
    cin = 2 * Cf + 10 + nf
    cout = 64
    w = (torch.rand(cout, cin, ksize, ksize, generator=g) - 0.5).half().float()
    packed, c0, c1, n_cta, nsplit = _pack(L, 1, Cf, nf, ksize, cout, w)
    assert c0 == 2 * Cf + 64 and c1 == nf
    wd = _decode(packed, ksize, c0 + c1, cout, n_cta, nsplit)
    # a reference-order input and the same data in our layout
    img0, img1 = torch.rand(1, 3, 8, 6, generator=g), torch.rand(1, 3, 8, 6, generator=g)
    f0, f1 = torch.rand(1, Cf, 8, 6, generator=g), torch.rand(1, Cf, 8, 6, generator=g)
    bwd, fwd = torch.rand(1, 2, 8, 6, generator=g), torch.rand(1, 2, 8, 6, generator=g)
    net = torch.rand(1, nf, 8, 6, generator=g)
    ref_in = torch.cat([img0, f0, img1, f1, bwd, fwd, net], 1)
    misc = torch.cat([img0, img1, bwd, fwd, torch.full((1, 54, 8, 6), 7.0)], 1)   # channels 10..63 carry zero weights
    ours = torch.cat([f0, f1, misc, net], 1)
    ref = F.conv2d(ref_in, w, padding="same")
    assert (_conv_same(ours, wd) - ref).abs().max().item() <= 1e-4
