"""Pin oracle/gmfss.py (GMFSS Fortuna union downstream of its flow network: MetricNet, FeatureNet, GridNet, the splatting /
RIFE / fusion glue of Model.inference) and oracle/gmflow.py (GMFlow) to outputs of the unmodified reference model
(tests/golden/gmfss_*.npz, made by tools/make_golden_gmfss.py): the downstream part from the golden flows, GMFlow on its own,
and the whole model from the two frames."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_gmfss import gmfss_cases  # noqa: E402
from oracle import film as OF  # noqa: E402
from oracle import gmfss as OG  # noqa: E402
from oracle import gmfss_weights as GW  # noqa: E402


@pytest.mark.parametrize("name", sorted(gmfss_cases().keys()))
def test_gmfss_downstream_of_gmflow_matches_reference(name):
    cfg = gmfss_cases()[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sds = GW.synthetic_state_dicts(cfg["seed"])
    fr = OF.synthetic_clip(2, cfg["h"], cfg["w"], seed=cfg["clip_seed"]).permute(0, 3, 1, 2).contiguous()
    h, w = cfg["h"], cfg["w"]
    ph, pw = ((h - 1) // 64 + 1) * 64, ((w - 1) // 64 + 1) * 64      # CommonModelInference.forward, gmfss_fortuna/__init__.py:41-47
    i0, i1 = F.pad(fr[0:1], (0, pw - w, 0, ph - h)), F.pad(fr[1:2], (0, pw - w, 0, ph - h))
    f01, f10 = torch.from_numpy(g["flow01"]), torch.from_numpy(g["flow10"])
    with torch.no_grad():
        m0, m1, f1, f2 = OG.reuse_from_flows(sds, i0, i1, f01, f10)
        assert (m0 - torch.from_numpy(g["metric0"])).abs().max().item() <= 1e-4
        assert (m1 - torch.from_numpy(g["metric1"])).abs().max().item() <= 1e-4
        out = OG.inference(sds, i0, i1, f01, f10, m0, m1, f1, f2, cfg["t"])[:, :, :h, :w]
    assert (out - torch.from_numpy(g["out"])).abs().max().item() <= 2e-4


@pytest.mark.parametrize("name", sorted(gmfss_cases().keys()))
def test_gmflow_matches_reference_flows(name):
    """oracle/gmflow.py (encoder, swin transformer, global / local matching, propagation, convex up-sampling) against the two
    flows the unmodified reference's GMFlow produced inside Model.reuse (half-size frames)."""
    from oracle import gmflow as GF
    cfg = gmfss_cases()[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sd = GW.synthetic_state_dicts(cfg["seed"])["flownet"]
    fr = OF.synthetic_clip(2, cfg["h"], cfg["w"], seed=cfg["clip_seed"]).permute(0, 3, 1, 2).contiguous()
    h, w = cfg["h"], cfg["w"]
    ph, pw = ((h - 1) // 64 + 1) * 64, ((w - 1) // 64 + 1) * 64
    i0, i1 = F.pad(fr[0:1], (0, pw - w, 0, ph - h)), F.pad(fr[1:2], (0, pw - w, 0, ph - h))
    h0 = F.interpolate(i0, scale_factor=0.5, mode="bilinear", align_corners=False)
    h1 = F.interpolate(i1, scale_factor=0.5, mode="bilinear", align_corners=False)
    with torch.no_grad():
        f01, f10 = GF.gmflow(sd, h0, h1), GF.gmflow(sd, h1, h0)
    assert f01.shape == g["flow01"].shape
    assert (f01 - torch.from_numpy(g["flow01"])).abs().max().item() <= 1e-3        # pixels
    assert (f10 - torch.from_numpy(g["flow10"])).abs().max().item() <= 1e-3


@pytest.mark.parametrize("name", sorted(gmfss_cases().keys()))
def test_gmfss_whole_model_matches_reference(name):
    cfg = gmfss_cases()[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sds = GW.synthetic_state_dicts(cfg["seed"])
    fr = OF.synthetic_clip(2, cfg["h"], cfg["w"], seed=cfg["clip_seed"]).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        out = OG.interpolate(sds, fr[0:1], fr[1:2], cfg["t"])
    assert (out - torch.from_numpy(g["out"])).abs().max().item() <= 5e-4
