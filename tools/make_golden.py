#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run from the repo root (build container only; /root/reference does not exist on the GPU box):

    python tools/make_golden.py

How: a stub ``comfy.model_management`` (device = cpu) is put on sys.path together with
/root/reference; ``vfi_models.rife.rife_arch.IFNet("4.6")`` and ``vfi_models.rife.RIFE_VFI``
are imported as they are; the checkpoint table gets ``"rife46.pth": "4.6"`` (absent at this
commit) and the downloader is replaced by a function returning a temp file with
``oracle.rife46.synthetic_state_dict(seed)`` (no weights ship with the reference, no network).
Inputs are regenerated from seeds by the tests, so only the reference OUTPUTS are stored.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("VFI_REFERENCE", "/root/reference")

from oracle import rife46 as O  # noqa: E402


def _install_stub():
    comfy = types.ModuleType("comfy")
    mm = types.ModuleType("comfy.model_management")
    mm.get_torch_device = lambda: torch.device("cpu")
    mm.soft_empty_cache = lambda *a, **k: None
    mm.is_nvidia = lambda: False
    mm.get_torch_device_name = lambda d: str(d)
    comfy.model_management = mm
    sys.modules["comfy"] = comfy
    sys.modules["comfy.model_management"] = mm
    sys.path.insert(0, REF)


def cases():
    """Shared with tests/test_oracle_golden.py: name -> kwargs."""
    return {
        # IFNet level: pad 96x160 -> 128x192, two timesteps in one batch
        "ifnet_96x160": dict(kind="ifnet", seed=0, gain=1.0, h=96, w=160, ts=(0.5, 0.25), clip_seed=11),
        # larger flows (lastconv x4) so border clamping and big displacements are exercised
        "ifnet_64x64_gain4": dict(kind="ifnet", seed=3, gain=4.0, h=64, w=64, ts=(0.5,), clip_seed=12),
        # already a multiple of 64: no padding
        "ifnet_128x128_rand": dict(kind="ifnet", seed=5, gain=1.0, h=128, w=128, ts=(0.75,), clip_seed=-1),
        # node level: multiplier 3 with a skip list, 4-channel input (alpha dropped)
        "node_m3_skip": dict(kind="node", seed=1, gain=1.0, n=4, h=40, w=72, c=4, multiplier=3,
                             states=([1], True), clip_seed=13),
        # node level: per-pair multiplier list shorter than the pair count (padded with 2), entry 1 => no mids
        "node_mlist": dict(kind="node", seed=2, gain=2.0, n=5, h=64, w=64, c=3, multiplier=[2, 1, 3],
                           states=None, clip_seed=14),
        # arch 4.7 (rife47.pth / rife49.pth): encode head, feature warps, replaced mask
        "ifnet47_96x160": dict(kind="ifnet", arch="4.7", seed=8, gain=1.0, h=96, w=160, ts=(0.5, 0.3), clip_seed=16),
        "ifnet47_64x128_gain3": dict(kind="ifnet", arch="4.7", seed=9, gain=3.0, h=64, w=128, ts=(0.5,), clip_seed=17),
        "node47_m2": dict(kind="node", arch="4.7", ckpt="rife49.pth", seed=10, gain=1.0, n=3, h=56, w=88, c=3,
                          multiplier=2, states=None, clip_seed=18),
        # arch 4.17 (rife417.pth): Head_417 encoder (three 32-channel convs + ConvT), 8 feature channels per frame
        "ifnet417_96x160": dict(kind="ifnet", arch="4.17", seed=20, gain=1.0, h=96, w=160, ts=(0.5, 0.7), clip_seed=21),
        "ifnet417_64x128_gain3": dict(kind="ifnet", arch="4.17", seed=22, gain=3.0, h=64, w=128, ts=(0.25,),
                                      clip_seed=23),
        "node417_m3": dict(kind="node", arch="4.17", ckpt="rife417.pth", seed=24, gain=1.0, n=3, h=56, w=88, c=3,
                           multiplier=3, states=None, clip_seed=25),
        # arch 4.26 (rife426.pth): five blocks (scales 16..1), Head encoder, 13-channel lastconv with feature feedback
        "ifnet426_96x160": dict(kind="ifnet", arch="4.26", seed=30, gain=1.0, h=96, w=160, ts=(0.5, 0.35), clip_seed=31),
        "ifnet426_128x128_gain3": dict(kind="ifnet", arch="4.26", seed=32, gain=3.0, h=128, w=128, ts=(0.75,),
                                       clip_seed=33),
        "node426_m2": dict(kind="node", arch="4.26", ckpt="rife426.pth", seed=34, gain=1.0, n=3, h=56, w=88, c=3,
                           multiplier=2, states=None, clip_seed=35),
        # node scale_factor 2 / 4 (rife/__init__.py:156-160: scale_list / scale_factor): the last one / two blocks run at
        # scale 0.5 / 0.25, i.e. on an UP-scaled input
        "ifnet_64x64_sf2": dict(kind="ifnet", seed=40, gain=2.0, h=64, w=64, ts=(0.5,), clip_seed=41, scale_factor=2.0),
        # not square, padded (40x100 -> 64x128): rows / columns of the up-scaled blocks must not be interchangeable
        "ifnet_40x100_sf2": dict(kind="ifnet", seed=54, gain=2.0, h=40, w=100, ts=(0.35,), clip_seed=55, scale_factor=2.0),
        "ifnet426_40x100_sf4": dict(kind="ifnet", arch="4.26", seed=56, gain=2.0, h=40, w=100, ts=(0.5,), clip_seed=57,
                                    scale_factor=4.0),
        "ifnet_64x64_sf4": dict(kind="ifnet", seed=42, gain=2.0, h=64, w=64, ts=(0.4,), clip_seed=43, scale_factor=4.0),
        "ifnet47_64x64_sf2": dict(kind="ifnet", arch="4.7", seed=44, gain=2.0, h=64, w=64, ts=(0.5,), clip_seed=45, scale_factor=2.0),
        "ifnet47_64x64_sf4": dict(kind="ifnet", arch="4.7", seed=46, gain=2.0, h=64, w=64, ts=(0.6,), clip_seed=47,
                                   scale_factor=4.0),
        "ifnet417_64x64_sf2": dict(kind="ifnet", arch="4.17", seed=48, gain=2.0, h=64, w=64, ts=(0.5,), clip_seed=49,
                                   scale_factor=2.0),
        "ifnet426_64x64_sf2": dict(kind="ifnet", arch="4.26", seed=50, gain=2.0, h=64, w=64, ts=(0.5,), clip_seed=51,
                                   scale_factor=2.0),
        "ifnet426_64x64_sf4": dict(kind="ifnet", arch="4.26", seed=52, gain=2.0, h=64, w=64, ts=(0.45,), clip_seed=53,
                                   scale_factor=4.0),
        # node level: keep-list (is_skip_list False)
        "node_keep": dict(kind="node", seed=4, gain=1.0, n=4, h=48, w=80, c=3, multiplier=2,
                          states=([0, 2], False), clip_seed=15),
    }


def loop_cases():
    """generic_frame_loop cases (vfi_utils.py:339-389) with a stand-in model function; shared with the tests."""
    return {
        "loop_ts_m3_skip": dict(n=5, multiplier=3, states=([1, 3], True), use_timestep=True),
        "loop_bisect_m4": dict(n=3, multiplier=4, states=None, use_timestep=False),
        "loop_bisect_m7_keep": dict(n=4, multiplier=7, states=([0, 2], False), use_timestep=False),
        "loop_list": dict(n=5, multiplier=[3, 0, 2], states=([0], False), use_timestep=True),
        "loop_list_bisect": dict(n=4, multiplier=[2, 5], states=None, use_timestep=False),
    }


def loop_model(f0, f1, t, gain):
    """Stand-in for a model's middle-frame function: order sensitive, timestep sensitive."""
    if t is None:
        return 0.5 * (f0 + f1) + gain * (f1 - f0).abs()
    return (1 - t) * f0 + t * f1 + gain * t * t


def loop_frames(n):
    g = torch.Generator().manual_seed(100 + n)
    return torch.rand(n, 3, 6, 8, generator=g)


def make_inputs(cfg):
    if cfg["kind"] == "ifnet":
        if cfg["clip_seed"] < 0:
            g = torch.Generator().manual_seed(99)
            fr = torch.rand(2, cfg["h"], cfg["w"], 3, generator=g) * 1.2 - 0.1  # also exercises the clamp
        else:
            fr = O.synthetic_clip(2, cfg["h"], cfg["w"], seed=cfg["clip_seed"])
        return fr
    fr = O.synthetic_clip(cfg["n"], cfg["h"], cfg["w"], seed=cfg["clip_seed"])
    if cfg["c"] == 4:
        fr = torch.cat([fr, torch.ones_like(fr[..., :1])], -1)
    return fr


def main():
    _install_stub()
    import vfi_models.rife as R
    from vfi_models.rife.rife_arch import IFNet
    from vfi_utils import InterpolationStateList

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    R.CKPT_NAME_VER_DICT["rife46.pth"] = "4.6"
    only = set(sys.argv[1:])  # optional: regenerate only the named cases
    for name, cfg in cases().items():
        if only and name not in only:
            continue
        arch = cfg.get("arch", "4.6")
        sd = O.synthetic_state_dict(cfg["seed"], cfg["gain"], arch=arch)
        fr = make_inputs(cfg)
        if cfg["kind"] == "ifnet":
            m = IFNet(arch_ver=arch).eval()
            m.load_state_dict(sd)
            x = fr.permute(0, 3, 1, 2)
            ts = torch.tensor(cfg["ts"], dtype=torch.float32).view(-1, 1, 1, 1)
            b = len(cfg["ts"])
            with torch.inference_mode():
                sl = [v / cfg.get("scale_factor", 1.0) for v in O.SCALE_LIST[arch]]
                out = m(x[0:1].repeat(b, 1, 1, 1), x[1:2].repeat(b, 1, 1, 1), ts, sl, False, False)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), out=out.numpy())
        else:
            with tempfile.TemporaryDirectory() as td:
                ckpt = cfg.get("ckpt", "rife46.pth")
                path = os.path.join(td, ckpt)
                torch.save(sd, path)
                R.load_file_from_github_release = lambda model_type, ckpt_name, _p=path: _p
                R._model_cache.clear()
                st = None
                if cfg["states"] is not None:
                    st = InterpolationStateList(list(cfg["states"][0]), cfg["states"][1])
                (out,) = R.RIFE_VFI().vfi(ckpt, fr, multiplier=cfg["multiplier"],
                                          optional_interpolation_states=st)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), out=out.numpy())
        print(name, tuple(out.shape), float(out.mean()))
    import vfi_utils as VU
    for name, cfg in loop_cases().items():
        if only and name not in only:
            continue
        st = None if cfg["states"] is None else InterpolationStateList(list(cfg["states"][0]), cfg["states"][1])
        out = VU.generic_frame_loop("Stand_In_VFI", loop_frames(cfg["n"]), 10, cfg["multiplier"], loop_model, 0.03,
                                    interpolation_states=st, use_timestep=cfg["use_timestep"], dtype=torch.float32)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), out=out.numpy())
        print(name, tuple(out.shape), float(out.mean()))


if __name__ == "__main__":
    main()
