"""CPU: the op restatements (oracle/ops_ref.py) against independent formulations."""
import torch
import torch.nn.functional as F

from oracle import ops_ref as R


def test_softsplat_is_adjoint_of_backward_warp():
    """<splat(a, flow), b> == <a, backwarp_zeros(b, flow)> (summation splat = adjoint of bilinear sampling with zero
    padding)."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 3, 12, 15, generator=g)
    b = torch.randn(2, 3, 12, 15, generator=g)
    flow = 3 * torch.randn(2, 2, 12, 15, generator=g)
    s = R.softsplat_sum(a, flow)
    ys, xs = torch.meshgrid(torch.arange(12.), torch.arange(15.), indexing="ij")
    gx = (xs + flow[:, 0]) / 14 * 2 - 1
    gy = (ys + flow[:, 1]) / 11 * 2 - 1
    w = F.grid_sample(b, torch.stack([gx, gy], -1), mode="bilinear", padding_mode="zeros", align_corners=True)
    assert abs((s * b).sum().item() - (a * w).sum().item()) < 1e-3


def test_softsplat_tiny_loops_and_nonfinite():
    a = torch.arange(12.).reshape(1, 1, 3, 4) + 1
    flow = torch.zeros(1, 2, 3, 4)
    flow[0, 0, 1, 1] = 0.25      # 75 % stays, 25 % moves right
    flow[0, 1, 2, 3] = 5.0       # leaves the image
    flow[0, 0, 0, 0] = float("nan")
    s = R.softsplat_sum(a, flow)
    exp = a.clone()
    exp[0, 0, 1, 1] = a[0, 0, 1, 1] * 0.75
    exp[0, 0, 1, 2] = a[0, 0, 1, 2] + a[0, 0, 1, 1] * 0.25
    exp[0, 0, 2, 3] = 0
    exp[0, 0, 0, 0] = 0
    assert torch.allclose(s, exp)
    m = torch.zeros(1, 1, 3, 4)
    soft = R.softsplat(a, flow, m, "soft")
    assert torch.isfinite(soft).all()


def test_costvol_and_correlation_vs_unfold():
    g = torch.Generator().manual_seed(1)
    one, two = torch.randn(2, 5, 9, 11, generator=g), torch.randn(2, 5, 9, 11, generator=g)
    cv = R.costvol_l1(one, two)
    cr = R.correlation_dot(one, two)
    patches = F.unfold(F.pad(two, (4, 4, 4, 4)), 9).reshape(2, 5, 81, 9, 11)
    assert torch.allclose(cr, (one[:, :, None] * patches).mean(1), atol=1e-6)
    inside = F.unfold(F.pad(torch.ones(2, 1, 9, 11), (4, 4, 4, 4)), 9).reshape(2, 1, 81, 9, 11)
    l1 = ((one[:, :, None] - patches * inside).abs()).mean(1)
    assert torch.allclose(cv, l1, atol=1e-6)


def test_sepconv_vs_explicit():
    g = torch.Generator().manual_seed(2)
    k, h, w = 5, 4, 6
    x = torch.randn(1, 2, h + k - 1, w + k - 1, generator=g)
    ver, hor = torch.randn(1, k, h, w, generator=g), torch.randn(1, k, h, w, generator=g)
    out = R.sepconv(x, ver, hor)
    ref = torch.zeros(1, 2, h, w)
    for y in range(h):
        for xx in range(w):
            kern = ver[0, :, y, xx, None] * hor[0, None, :, y, xx]
            ref[0, :, y, xx] = (x[0, :, y:y + k, xx:xx + k] * kern).sum((1, 2))
    assert torch.allclose(out, ref, atol=1e-5)
