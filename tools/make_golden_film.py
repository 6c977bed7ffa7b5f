#!/usr/bin/env python
"""Generate tests/golden/film_*.npz by running the UNMODIFIED reference FILM code in this container.

    python tools/make_golden_film.py [case ...]

``vfi_models.film.film_arch.Interpolator`` is imported as it is and loaded with
``oracle.film.synthetic_state_dict(seed)``.  For the node cases the reference's ``FILM_VFI.vfi`` runs unmodified:
its checkpoint is a TorchScript file (``torch.jit.load``, film/__init__.py:74), so the unmodified Interpolator is
scripted with ``torch.jit.script`` and saved to a temp file that the stubbed downloader returns.
Inputs are regenerated from seeds by the tests; only the reference OUTPUTS are stored (fp16-rounded copies of the
intermediate flows are not needed: the per-level flows are small and stored in fp32).
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from make_golden import _install_stub  # noqa: E402
from oracle import film as OF  # noqa: E402


def film_cases():
    """Shared with tests/test_oracle_film.py and the GPU tests: name -> kwargs."""
    return {
        # model level; 128x192 so the coarsest of the 7 pyramid levels is 2x3
        "film_net_128x192": dict(kind="net", seed=0, flow_gain=1.0, h=128, w=192, clip_seed=41),
        # sizes that are NOT multiples of 64 (the node does not pad): odd sizes appear down the pyramid
        "film_net_72x104": dict(kind="net", seed=1, flow_gain=2.0, h=72, w=104, clip_seed=42),
        # white-noise frames outside [0,1]: nothing clamps the inputs in the model
        "film_net_64x64_rand": dict(kind="net", seed=2, flow_gain=1.0, h=64, w=64, clip_seed=-1),
        # node level: multiplier 4 (3 recursive calls per pair), a skip list (the pair is dropped with its frame)
        "film_node_m4_skip": dict(kind="node", seed=3, flow_gain=1.0, n=4, h=64, w=96, c=4, multiplier=4,
                                  states=([1], True), clip_seed=43),
        # node level: multiplier list shorter than the pair count (padded with 2)
        "film_node_mlist": dict(kind="node", seed=4, flow_gain=1.0, n=4, h=64, w=64, c=3, multiplier=[3, 2],
                                states=None, clip_seed=44),
    }


def film_inputs(cfg):
    if cfg["kind"] == "net":
        if cfg["clip_seed"] < 0:
            g = torch.Generator().manual_seed(98)
            return torch.rand(2, cfg["h"], cfg["w"], 3, generator=g) * 1.2 - 0.1
        return OF.synthetic_clip(2, cfg["h"], cfg["w"], seed=cfg["clip_seed"])
    fr = OF.synthetic_clip(cfg["n"], cfg["h"], cfg["w"], seed=cfg["clip_seed"])
    if cfg["c"] == 4:
        fr = torch.cat([fr, torch.ones_like(fr[..., :1])], -1)
    return fr


def main():
    _install_stub()
    import vfi_models.film as FM
    from vfi_models.film.film_arch import Interpolator
    from vfi_utils import InterpolationStateList

    out_dir = os.path.join(ROOT, "tests", "golden")
    only = set(sys.argv[1:])
    for name, cfg in film_cases().items():
        if only and name not in only:
            continue
        sd = OF.synthetic_state_dict(cfg["seed"], cfg["flow_gain"])
        m = Interpolator().eval()
        m.load_state_dict(sd)
        fr = film_inputs(cfg)
        if cfg["kind"] == "net":
            x = fr.permute(0, 3, 1, 2)
            with torch.no_grad():
                d = m.debug_forward(x[0:1], x[1:2], torch.full((1, 1), 0.5))
            out = d["image"][0]
            extra = {f"fwd_flow{l}": d["forward_flow_pyramid"][l].numpy() for l in range(5)}
            extra.update({f"bwd_flow{l}": d["backward_flow_pyramid"][l].numpy() for l in range(5)})
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), out=out.numpy(), **extra)
            print(name, tuple(out.shape), float(out.mean()), float(out.std()), float(out.min()), float(out.max()),
                  "flow0 absmax", float(d["forward_flow_pyramid"][0].abs().max()))
        else:
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "film_net_fp32.pt")
                torch.jit.script(m).save(path)
                FM.load_file_from_github_release = lambda model_type, ckpt_name, _p=path: _p
                st = None
                if cfg["states"] is not None:
                    st = InterpolationStateList(list(cfg["states"][0]), cfg["states"][1])
                (out,) = FM.FILM_VFI().vfi("film_net_fp32.pt", fr, multiplier=cfg["multiplier"],
                                           optional_interpolation_states=st)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), out=out.numpy())
            print(name, tuple(out.shape), float(out.mean()))


if __name__ == "__main__":
    main()
