"""-m gpu: the vfi_models/ops replacements (softsplat / costvol / correlation / sepconv) through the C ABI against the
CPU restatement of the reference kernels.  fp32 everywhere; tolerance covers summation order only
(atomics in the splat, Kahan vs float64 in sepconv)."""
import pytest
import torch

from oracle import ops_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(pkg):
    import cfi_b200.ops as O
    return O


def _close(a, b, tol=2e-5):
    return (a.cpu() - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def test_softsplat_modes(ops):
    g = torch.Generator().manual_seed(0)
    for c, h, w in ((3, 37, 53), (65, 24, 40)):
        x = torch.randn(2, c, h, w, generator=g)
        flow = 6 * torch.randn(2, 2, h, w, generator=g)
        flow[0, 0, 3, 4] = float("nan")
        flow[1, 1, 5, 6] = float("inf")
        metric = 0.5 * torch.randn(2, 1, h, w, generator=g)
        assert _close(ops.softsplat_func.apply(x.cuda(), flow.cuda()), R.softsplat_sum(x, flow))
        for mode in ("sum", "avg", "linear", "soft", "soft-zeroeps", "linear-clipeps"):
            m = None if mode in ("sum", "avg") else (metric.abs() + 0.1 if "linear" in mode else metric)
            got = ops.softsplat(x.cuda(), flow.cuda(), None if m is None else m.cuda(), mode)
            assert _close(got, R.softsplat(x, flow, m, mode), 1e-4), mode


def test_costvol_and_correlation(ops):
    g = torch.Generator().manual_seed(1)
    for c, h, w in ((32, 20, 28), (7, 5, 3), (128, 12, 16)):
        a, b = torch.randn(2, c, h, w, generator=g), torch.randn(2, c, h, w, generator=g)
        assert _close(ops.costvol_func.apply(a.cuda(), b.cuda()), R.costvol_l1(a, b))
        assert _close(ops.FunctionCorrelation(a.cuda(), b.cuda()), R.correlation_dot(a, b))


def test_sepconv(ops):
    g = torch.Generator().manual_seed(2)
    for c, k, h, w in ((4, 51, 20, 24), (3, 5, 9, 7), (6, 13, 8, 8)):
        x = torch.rand(2, c, h + k - 1, w + k - 1, generator=g)
        ver = torch.randn(2, k, h, w, generator=g) / k
        hor = torch.randn(2, k, h, w, generator=g) / k
        assert _close(ops.sepconv_func.apply(x.cuda(), ver.cuda(), hor.cuda()), R.sepconv(x, ver, hor), 1e-5)


def test_cpu_tensors_raise(ops):
    from cfi_b200._lib import VfiError
    with pytest.raises(VfiError):
        ops.softsplat_func.apply(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4))
