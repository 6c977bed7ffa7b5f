#!/usr/bin/env python
"""Goldens at the sizes BASELINE.json quotes its numbers on, from the UNMODIFIED reference (build container only):

    python tools/make_golden_configs.py

* ``cfg1_anime_540p.npz`` - config 1: the reference's own demo pair ``demo_frames/anime0.png`` + ``anime1.png``
  (540 x 960, padded to 576 x 960 inside the model) through the whole, unmodified ``RIFE_VFI.vfi`` (arch 4.6, 2x, seeded
  synthetic weights - none ship with the reference).  The two decoded input frames are stored as uint8 (the GPU box has
  no /root/reference), the interpolated frame as 16-bit fixed point (quantisation floor 107 dB).
* ``cfg2_1080p_arch{46,47}.npz`` - the 1080p geometry of config 2 (1080 x 1920 padded to 1088 x 1920): one pair of the
  bench's synthetic clip through the unmodified ``IFNet`` (arch 4.6 and 4.7, the node's default checkpoint family), two
  timesteps.  A full 1080p frame is 25 MB in fp32, so the fixture keeps (a) twelve 128 x 128 crops - the four corners,
  the four edge centres, the centre and three seeded interior positions - in 16-bit fixed point, and (b) the 8 x 8
  box-filtered whole frame, also in 16-bit fixed point, which catches a regional error anywhere in the frame.  The oracle is held to these
  on the CPU (tests/test_oracle_configs.py); the GPU tests compare the CUDA path with the crops, the box-filtered frame
  and the oracle's full frame (tests/test_gpu_configs.py).
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import REF, _install_stub  # noqa: E402

from oracle import rife46 as O  # noqa: E402

CROP = 128
CFG2 = dict(h=1080, w=1920, clip_seed=1234, ts=(0.5, 0.25), weight_seed=0, gain=1.0)
CFG1 = dict(weight_seed=0, gain=1.0, multiplier=2)


def crop_origins(h, w):
    """(y, x) of the twelve crops: corners, edge centres, centre, three seeded interior positions."""
    ys, xs = [0, (h - CROP) // 2, h - CROP], [0, (w - CROP) // 2, w - CROP]
    pts = [(y, x) for y in ys for x in xs]
    rng = np.random.RandomState(7)
    pts += [(int(rng.randint(0, h - CROP)), int(rng.randint(0, w - CROP))) for _ in range(3)]
    return pts


def box8(x):
    """[B, H, W, 3] -> [B, H/8, W/8, 3] mean over 8 x 8 boxes (H, W multiples of 8)."""
    b, h, w, c = x.shape
    return x.reshape(b, h // 8, 8, w // 8, 8, c).astype(np.float64).mean(axis=(2, 4)).astype(np.float32)


def q16(x):
    return np.round(np.clip(x, 0.0, 1.0) * 65535.0).astype(np.uint16)


def main():
    _install_stub()
    import vfi_models.rife as R
    from vfi_models.rife.rife_arch import IFNet
    from PIL import Image

    out_dir = os.path.join(ROOT, "tests", "golden")
    torch.set_num_threads(os.cpu_count())

    # ---- config 1: the demo pair through the whole node
    f0 = np.asarray(Image.open(os.path.join(REF, "demo_frames", "anime0.png")).convert("RGB"))
    f1 = np.asarray(Image.open(os.path.join(REF, "demo_frames", "anime1.png")).convert("RGB"))
    frames_u8 = np.stack([f0, f1])
    fr = torch.from_numpy(frames_u8).float() / 255.0  # ComfyUI's LoadImage: uint8 / 255 -> fp32 NHWC
    R.CKPT_NAME_VER_DICT["rife46.pth"] = "4.6"
    sd = O.synthetic_state_dict(CFG1["weight_seed"], CFG1["gain"], arch="4.6")
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "rife46.pth")
        torch.save(sd, path)
        R.load_file_from_github_release = lambda model_type, ckpt_name, _p=path: _p
        R._model_cache.clear()
        (out,) = R.RIFE_VFI().vfi("rife46.pth", fr, multiplier=CFG1["multiplier"])
    assert out.shape == (3, 540, 960, 3) and torch.equal(out[0], fr[0]) and torch.equal(out[2], fr[1])
    np.savez_compressed(os.path.join(out_dir, "cfg1_anime_540p.npz"), frames_u8=frames_u8, mid_q16=q16(out[1].numpy()))
    print("cfg1_anime_540p", tuple(out.shape), float(out[1].mean()))

    # ---- config 2 geometry: one 1080p pair of the bench clip through the unmodified IFNet
    fr = O.synthetic_clip(2, CFG2["h"], CFG2["w"], seed=CFG2["clip_seed"])
    x = fr.permute(0, 3, 1, 2)
    ts = torch.tensor(CFG2["ts"], dtype=torch.float32).view(-1, 1, 1, 1)
    b = len(CFG2["ts"])
    for arch in ("4.6", "4.7"):
        m = IFNet(arch_ver=arch).eval()
        m.load_state_dict(O.synthetic_state_dict(CFG2["weight_seed"], CFG2["gain"], arch=arch))
        with torch.inference_mode():
            out = m(x[0:1].repeat(b, 1, 1, 1), x[1:2].repeat(b, 1, 1, 1), ts, list(O.SCALE_LIST[arch]), False, False)
        out = out.clamp(0, 1).permute(0, 2, 3, 1).numpy()  # the node's clamp (rife/__init__.py:207), NHWC
        crops = np.stack([out[:, y:y + CROP, x0:x0 + CROP] for y, x0 in crop_origins(CFG2["h"], CFG2["w"])], 1)
        name = "cfg2_1080p_arch" + arch.replace(".", "")
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), crops_q16=q16(crops), box8_q16=q16(box8(out)))
        print(name, out.shape, float(out.mean()))


if __name__ == "__main__":
    main()
