"""Seeded stand-in weights for the five sub-networks of GMFSS Fortuna (union) - SURVEY.md section 8 row a11.
TEST INFRASTRUCTURE ONLY.

No GMFSS checkpoint ships with the reference and there is no network, so the golden vectors of
``tools/make_golden_gmfss.py`` (outputs of the UNMODIFIED ``GMFSS_Fortuna_union_arch.Model``) are made on weights
generated here from ``tests/golden/gmfss_spec.json`` - the names and shapes of the five ``state_dict()``s as the reference
builds them (``flownet`` = GMFlow 4.72 M, ``ifnet`` = IFNet("4.6") 5.31 M, ``metricnet`` 0.12 M, ``feat_ext`` 0.81 M,
``fusionnet`` = GridNet 7.84 M parameters).  The recipe does not import the reference, so a GPU box can regenerate the
same tensors: variance-preserving uniform weights for conv / linear tensors, PReLU slopes 0.25 +- 0.1, norm scales near
1, small biases.  GMFSS itself is NOT built in this repo yet (DESIGN.md coverage table); this file and the goldens pin the
target for it.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPEC = os.path.join(ROOT, "tests", "golden", "gmfss_spec.json")
NETS = ("flownet", "ifnet", "metricnet", "feat_ext", "fusionnet")


def load_spec() -> Dict[str, list]:
    with open(SPEC) as fh:
        return json.load(fh)


def synthetic_state_dicts(seed: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    spec = load_spec()
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, net in enumerate(NETS):
        g = torch.Generator().manual_seed(3000 + 10 * seed + k)
        sd: Dict[str, torch.Tensor] = {}
        for name, shape, dtype in spec[net]:
            shape = tuple(shape)
            if dtype != "torch.float32":                    # e.g. BatchNorm's num_batches_tracked
                sd[name] = torch.zeros(shape, dtype=getattr(torch, dtype.split(".")[1]))
                continue
            n = 1
            for s in shape:
                n *= s
            if len(shape) >= 2:                            # conv / linear / transposed conv weights
                fan_in = n // shape[0]
                v = (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / max(fan_in, 1)) ** 0.5
            elif name.endswith("running_var"):
                v = 1.0 + 0.1 * torch.rand(shape, generator=g)
            elif name.endswith("running_mean"):
                v = 0.05 * (torch.rand(shape, generator=g) * 2 - 1)
            elif name.endswith(".weight") and ("norm" in name or "bn" in name):
                v = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
            elif name.endswith(".weight") or name.endswith("beta"):  # PReLU slopes (1-D "weight"), ResConv beta
                v = (0.25 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)) if not name.endswith("beta") else \
                    1.0 + 0.25 * (torch.rand(shape, generator=g) * 2 - 1)
            else:                                          # biases and the like
                v = 0.05 * (torch.rand(shape, generator=g) * 2 - 1)
            sd[name] = v.float()
        out[net] = sd
    return out
