"""-m gpu: the whole RIFE-4.6 path through the C ABI / node against the oracle and the reference's golden outputs.

Tolerance (north_star): PSNR >= 50 dB of the interpolated frames vs the fp32 result, conv operands 16-bit,
everything else fp32.  Pass-through frames are bit exact.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import rife46 as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import cases, make_inputs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PSNR_MIN = 50.0


def _engine(pkg, sd, dtype="float32", batch=8):
    from cfi_b200.engine import Rife46Engine
    return Rife46Engine(sd, 0, dtype, batch=batch)


# (the scale_factor 2 / 4 cases have their own file, tests/test_gpu_zrife_scale.py: their two kernels were added after
# r01's last GPU minute)
@pytest.mark.parametrize("name", [n for n, c in cases().items() if c["kind"] == "ifnet" and "scale_factor" not in c])
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_ifnet_golden(pkg, name, dtype):
    """C-ABI forward vs the UNMODIFIED reference's output (tests/golden, made by tools/make_golden.py)."""
    cfg = cases()[name]
    ref = torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["out"]).clamp(0, 1)  # node clamps, :207
    sd = O.synthetic_state_dict(cfg["seed"], cfg["gain"], arch=cfg.get("arch", "4.6"))
    fr = make_inputs(cfg)
    eng = _engine(pkg, sd, dtype)
    b = len(cfg["ts"])
    out = eng.forward(fr.cuda().contiguous(), [0] * b, [1] * b, list(cfg["ts"])).cpu()
    eng.close()
    p = O.psnr(out, ref.permute(0, 2, 3, 1))
    print(f"{name} {dtype}: PSNR {p:.2f} dB")
    assert p >= PSNR_MIN


@pytest.mark.parametrize("arch", ["4.6", "4.7", "4.17", "4.26"])
def test_flow_and_mask_match_oracle(pkg, arch):
    """Intermediate state: final full-resolution flow / mask vs the oracle's (fp32)."""
    sd = O.synthetic_state_dict(0, arch=arch)
    fr = O.synthetic_clip(2, 96, 160, seed=11)
    eng = _engine(pkg, sd)
    eng.forward(fr.cuda(), [0], [1], [0.5])
    flow, mask = eng.debug_state(1)
    taps = {}
    O.ifnet_forward(arch, sd, fr[0:1].permute(0, 3, 1, 2), fr[1:2].permute(0, 3, 1, 2),
                    torch.tensor([0.5]).view(1, 1, 1, 1), taps=taps)
    eng.close()
    last = len(O.SCALE_LIST[arch]) - 1
    f_ref = taps[f"flow{last}"].permute(0, 2, 3, 1)
    m_ref = taps[f"mask{last}"][:, 0]
    ef = (flow.cpu() - f_ref).abs().max().item()
    em = (mask.cpu() - m_ref).abs().max().item()
    print(f"max |flow err| {ef:.4f} px (max |flow| {f_ref.abs().max():.2f}), max |mask err| {em:.4f}")
    assert ef < 0.05 and em < 0.05


@pytest.mark.parametrize("scale_factor,h,w", [(0.5, 128, 192), (0.25, 128, 256)])
def test_scale_factor(pkg, scale_factor, h, w):
    """The node's scale_factor widget rescales the block scales (rife/__init__.py:157-160): 16,8,4,2 / 32,16,8,4."""
    sd = O.synthetic_state_dict(6)
    fr = O.synthetic_clip(2, h, w, seed=21)
    eng = _engine(pkg, sd)
    out = eng.forward(fr.cuda(), [0], [1], [0.5], scale_factor=scale_factor).cpu()
    eng.close()
    sl = [8 / scale_factor, 4 / scale_factor, 2 / scale_factor, 1 / scale_factor]
    ref = O.ifnet46_forward(sd, fr[0:1].permute(0, 3, 1, 2), fr[1:2].permute(0, 3, 1, 2),
                            torch.tensor([0.5]).view(1, 1, 1, 1), scale_list=sl).clamp(0, 1).permute(0, 2, 3, 1)
    p = O.psnr(out, ref)
    print(f"scale_factor {scale_factor}: PSNR {p:.2f} dB")
    assert p >= PSNR_MIN


def test_scale_factor_426(pkg):
    """arch 4.26 at scale_factor 0.5: block scales 32,16,8,4,2 - the last front adds four coarser levels."""
    sd = O.synthetic_state_dict(26, arch="4.26")
    fr = O.synthetic_clip(2, 128, 256, seed=27)
    eng = _engine(pkg, sd)
    out = eng.forward(fr.cuda(), [0], [1], [0.4], scale_factor=0.5).cpu()
    eng.close()
    ref = O.ifnet426_forward(sd, fr[0:1].permute(0, 3, 1, 2), fr[1:2].permute(0, 3, 1, 2),
                             torch.tensor([0.4]).view(1, 1, 1, 1), scale_list=[32, 16, 8, 4, 2])
    p = O.psnr(out, ref.clamp(0, 1).permute(0, 2, 3, 1))
    print(f"arch 4.26 scale_factor 0.5: PSNR {p:.2f} dB")
    assert p >= PSNR_MIN


def test_batches_and_batch_size_do_not_change_results(pkg):
    sd = O.synthetic_state_dict(1)
    fr = O.synthetic_clip(4, 72, 104, seed=5).cuda()
    f0, f1, t = [0, 0, 1, 2, 2], [1, 1, 2, 3, 3], [0.25, 0.75, 0.5, 1 / 3, 2 / 3]
    outs = []
    for batch in (1, 2, 8):
        eng = _engine(pkg, sd, batch=batch)
        outs.append(eng.forward(fr, f0, f1, t).cpu())
        eng.close()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_host_pipeline_equals_device_path(pkg):
    sd = O.synthetic_state_dict(2)
    fr = O.synthetic_clip(6, 64, 96, seed=6)
    f0, f1, t = [0, 1, 2, 3, 4], [1, 2, 3, 4, 5], [0.5] * 5
    eng = _engine(pkg, sd, batch=2)
    dev = eng.forward(fr.cuda(), f0, f1, t).cpu()
    host = torch.empty(5, 64, 96, 3).pin_memory()
    eng.interpolate_host(fr.contiguous(), f0, f1, t, host)
    # slots + a frame sub-range (what one rank of a sharded run does)
    host2 = torch.zeros(9, 64, 96, 3)
    eng.interpolate_host(fr.contiguous(), f0[2:], f1[2:], t[2:], host2, out_slots=[8, 1, 4], frame_range=(2, 6))
    eng.close()
    assert torch.equal(dev, host)
    assert torch.equal(host2[8], dev[2]) and torch.equal(host2[1], dev[3]) and torch.equal(host2[4], dev[4])


@pytest.mark.parametrize("name", [n for n, c in cases().items() if c["kind"] == "node"])
def test_node_golden(pkg, name, tmp_path, monkeypatch):
    """The drop-in node vs the reference node's own output on the same frames / options."""
    import cfi_b200.node as N
    cfg = cases()[name]
    ref = torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["out"])
    sd = O.synthetic_state_dict(cfg["seed"], cfg["gain"], arch=cfg.get("arch", "4.6"))
    ckpt = cfg.get("ckpt", "rife46.pth")
    path = tmp_path / ckpt
    torch.save(sd, path)
    monkeypatch.setattr(N, "load_file_from_github_release", lambda model_type, ckpt_name: str(path))
    N._model_cache.clear()
    fr = make_inputs(cfg)
    st = None
    if cfg["states"] is not None:
        st = N.InterpolationStateList(list(cfg["states"][0]), cfg["states"][1])
    (out,) = N.RIFE_VFI().vfi(ckpt, fr, multiplier=cfg["multiplier"], optional_interpolation_states=st)
    N._model_cache.clear()
    assert out.shape == ref.shape and out.dtype == torch.float32 and not out.is_cuda
    tasks, mults = O.build_tasks(cfg["n"], cfg["multiplier"], cfg["states"])
    # which output slots are pass-through frames
    per_pair = [0] * (cfg["n"] - 1)
    for p, _ in tasks:
        per_pair[p] += 1
    slot, orig = 0, []
    for p in range(cfg["n"] - 1):
        orig.append(slot)
        slot += 1 + per_pair[p]
    orig.append(slot)
    for i, s in enumerate(orig):
        assert torch.equal(out[s], fr[i, ..., :3]), "pass-through frame must be bit exact"
    mids = [s for s in range(out.shape[0]) if s not in orig]
    p = O.psnr(out[mids], ref[mids])
    print(f"{name}: {len(mids)} interpolated frames, PSNR {p:.2f} dB")
    assert p >= PSNR_MIN


def test_errors_are_loud(pkg):
    from cfi_b200._lib import VfiError
    sd = O.synthetic_state_dict(0)
    eng = _engine(pkg, sd)
    fr = O.synthetic_clip(2, 64, 64, seed=1).cuda()
    with pytest.raises(VfiError):
        eng.forward(fr, [0], [5], [0.5])            # frame index out of range
    with pytest.raises(VfiError):
        eng.forward(fr, [0], [1], [0.5], scale_factor=3.0)   # not one of the widget's values 0.25 / 0.5 / 1 / 2 / 4
    eng.close()
