#!/usr/bin/env python
"""Generate tests/golden/sepconv_*.npz by running the UNMODIFIED reference Sepconv code in this container.

    python tools/make_golden_sepconv.py [case ...]

``vfi_models.sepconv.sepconv_enhanced.Network`` (and for the node case ``vfi_models.sepconv.SepconvVFI``) run as they
are on ``oracle.sepconv.synthetic_state_dict(seed)``.  The one thing that cannot be the reference's own is the custom op:
``vfi_models.ops`` imports cupy at module level (cupy_ops/utils.py:1), which is absent here, so ``sys.modules`` is
pre-seeded with a module whose ``sepconv_func.apply`` is ``oracle.ops_ref.sepconv`` (the CPU restatement of
cupy_ops/sepconv.py:86-117).  Inputs are regenerated from seeds by the tests; outputs and the four coefficient maps of
the model-level cases are stored.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from make_golden import _install_stub  # noqa: E402
from oracle import film as OF  # noqa: E402  (synthetic_clip)
from oracle import ops_ref  # noqa: E402
from oracle import sepconv as OS  # noqa: E402


def sepconv_cases():
    return {
        # all four encoder levels even (48 x 64 -> 24x32 -> 12x16 -> 6x8 -> 3x4)
        "sepconv_net_48x64": dict(kind="net", seed=0, h=48, w=64, clip_seed=51),
        # odd input size (replicate-padded to even), odd rows down the pyramid (the decoder's crop-by-one path)
        "sepconv_net_45x54": dict(kind="net", seed=1, h=45, w=54, clip_seed=52),
        # small odd case for the whole-path host emulation (tests/test_sepconv_full_host.py): 21x30 -> 22x30, rows 11x15, 6x8, 3x4, 2x2
        "sepconv_net_21x30": dict(kind="net", seed=3, h=21, w=30, clip_seed=54),
        # No node-level case: on this torch version (2.11) the unmodified SepconvVFI.vfi raises inside Network.forward
        # (`tenStack.view` at sepconv_enhanced.py:626 on the stack of the permuted frame views preprocess_frames hands it,
        # for even and odd sizes alike), so there is no reference node output to pin to.  The loop it would run,
        # generic_frame_loop(use_timestep=False), is pinned on its own (tests/golden/loop_bisect_*.npz).
    }


def sepconv_inputs(cfg):
    n = 2 if cfg["kind"] == "net" else cfg["n"]
    return OF.synthetic_clip(n, cfg["h"], cfg["w"], seed=cfg["clip_seed"])


def install_ops_stub():
    import vfi_models
    ops = types.ModuleType("vfi_models.ops")

    class _SepconvFunc:
        @staticmethod
        def apply(ten_in, ver, hor):
            return ops_ref.sepconv(ten_in, ver, hor)

    ops.sepconv_func = _SepconvFunc
    sys.modules["vfi_models.ops"] = ops
    vfi_models.ops = ops


def main():
    _install_stub()
    install_ops_stub()
    import vfi_models.sepconv as SM
    from vfi_models.sepconv.sepconv_enhanced import Network
    from vfi_utils import InterpolationStateList

    out_dir = os.path.join(ROOT, "tests", "golden")
    only = set(sys.argv[1:])
    for name, cfg in sepconv_cases().items():
        if only and name not in only:
            continue
        sd = OS.synthetic_state_dict(cfg["seed"])
        fr = sepconv_inputs(cfg)
        if cfg["kind"] == "net":
            m = Network().eval()
            m.load_state_dict(sd)
            x = fr.permute(0, 3, 1, 2).contiguous()   # Network.forward views the stacked pair (:626): contiguous NCHW
            out = m(x[0:1], x[1:2])
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), out=out.numpy())
            print(name, tuple(out.shape), float(out.mean()), float(out.std()), float(out.min()), float(out.max()))
        else:
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "sepconv.pth")
                torch.save(sd, path)
                SM.load_file_from_github_release = lambda model_type, ckpt_name, _p=path: _p
                st = None
                if cfg["states"] is not None:
                    st = InterpolationStateList(list(cfg["states"][0]), cfg["states"][1])
                (out,) = SM.SepconvVFI().vfi("sepconv.pth", fr, multiplier=cfg["multiplier"],
                                             optional_interpolation_states=st)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), out=out.numpy())
            print(name, tuple(out.shape), float(out.mean()))


if __name__ == "__main__":
    main()
