"""CPU restatement (fp32) of the reference's Sepconv ("revisiting sepconv") path - SURVEY.md section 8 row a12.
TEST INFRASTRUCTURE ONLY.

Follows ``vfi_models/sepconv/sepconv_enhanced.py`` (``Network`` :536-706 with its ``Basic`` / ``Encode`` / ``Decode``
building blocks :47-533) and the node ``vfi_models/sepconv/__init__.py:32-57`` (``generic_frame_loop`` with
``use_timestep=False``) of Fannovel16/ComfyUI-Frame-Interpolation @ 26545cc.  ATen operators are called functionally; the
one custom op, ``sepconv_func`` (``vfi_models/ops/cupy_ops/sepconv.py:86-117``), is ``oracle.ops_ref.sepconv``.

Pinning: ``tools/make_golden_sepconv.py`` runs the UNMODIFIED ``Network`` on seeded weights with ``vfi_models.ops``
pre-seeded by that op restatement (cupy cannot be imported here) and stores the outputs in
``tests/golden/sepconv_*.npz``; ``tests/test_oracle_sepconv.py`` holds this file to them.  So the network around the op
is pinned to the reference; the op itself is pinned only to its published loop (parity of the op: unpinned).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import ops_ref

CHANNELS = (32, 64, 128, 256, 512)  # Network.intChannels, sepconv_enhanced.py:541
KSIZE = 51                           # :598-612


def state_dict_spec() -> List[Tuple[str, Tuple[int, ...]]]:
    """Names / shapes of ``Network().state_dict()`` in the reference's order (88 tensors, 13,560,102 values)."""
    spec: List[Tuple[str, Tuple[int, ...]]] = [("netInput.weight", (16, 3, 3, 3)), ("netInput.bias", (16,))]

    def basic(prefix, idx_prelu0, convs):
        # convs: list of (module index, cin, cout); a PReLU (one scalar) precedes each conv except where idx is None
        for (pi, ci, cin, cout) in convs:
            if pi is not None:
                spec.append((f"{prefix}.netMain.{pi}.weight", (1,)))
            spec.append((f"{prefix}.netMain.{ci}.weight", (cout, cin, 3, 3)))
            spec.append((f"{prefix}.netMain.{ci}.bias", (cout,)))

    for r in range(1, 5):   # Encode: prelu-sconv-prelu-conv (:549-556)
        basic(f"netEncode.0.netVer.{r}", 0, [(0, 1, CHANNELS[r - 1], CHANNELS[r]), (2, 3, CHANNELS[r], CHANNELS[r])])
    for k, r in enumerate((4, 3, 2, 1)):   # Decode rows, coarse first: prelu-conv-prelu-conv+skip (:571-580)
        basic(f"netDecode.0.netHor.{k}", 0, [(0, 1, CHANNELS[r], CHANNELS[r]), (2, 3, CHANNELS[r], CHANNELS[r])])
    for k, r in enumerate((3, 2, 1), start=1):   # prelu-up-conv-prelu-conv
        basic(f"netDecode.0.netVer.{k}", 0, [(0, 2, CHANNELS[r + 1], CHANNELS[r]), (3, 4, CHANNELS[r], CHANNELS[r])])
    for head in ("netVerone", "netVertwo", "netHorone", "netHortwo"):   # up-conv-prelu-conv (:583-598)
        basic(head, None, [(None, 1, 64, 64), (2, 3, 64, KSIZE)])
    return spec


def synthetic_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded stand-in for ``sepconv.pth``: variance-preserving uniform conv weights, PReLU slopes in [0.1, 0.4], and head
    biases shaped as a positive bump around the kernel centre so the 51x51 kernels are blur-like and their sum (the
    normaliser, :697-699) stays away from zero, like a trained model's."""
    g = torch.Generator().manual_seed(2000 + seed)
    sd: Dict[str, torch.Tensor] = {}
    bump = torch.exp(-0.5 * ((torch.arange(KSIZE, dtype=torch.float32) - 25.0) / 3.0) ** 2)
    for name, shape in state_dict_spec():
        if shape == (1,):
            v = 0.1 + 0.3 * torch.rand(1, generator=g)
        elif name.endswith(".bias"):
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
            if shape == (KSIZE,):
                v = v * 0.2 + bump
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            v = (torch.rand(shape, generator=g) * 2 - 1) * (6.0 / (1.1 * fan_in)) ** 0.5
            if shape[0] == KSIZE:
                v = v * 0.02
        sd[name] = v.float()
    return sd


def _basic(sd, prefix: str, x: torch.Tensor, kind: str) -> torch.Tensor:
    """Basic.forward (:295-306) for the four layer strings the network uses."""
    w = lambda i: sd[f"{prefix}.netMain.{i}.weight"]   # noqa: E731
    b = lambda i: sd[f"{prefix}.netMain.{i}.bias"]     # noqa: E731
    if kind == "enc":      # prelu(0.25)-sconv(3)-prelu(0.25)-conv(3)
        y = F.conv2d(F.prelu(x, w(0)), w(1), b(1), stride=2, padding=1)
        return F.conv2d(F.prelu(y, w(2)), w(3), b(3), padding=1)
    if kind == "hor":      # prelu-conv-prelu-conv+skip (identity shortcut: equal channels, stride 1 - :222-223)
        y = F.conv2d(F.prelu(x, w(0)), w(1), b(1), padding=1)
        return F.conv2d(F.prelu(y, w(2)), w(3), b(3), padding=1) + x
    if kind == "dec":      # prelu-up(bilinear)-conv-prelu-conv
        y = F.interpolate(F.prelu(x, w(0)), scale_factor=2.0, mode="bilinear", align_corners=False)
        y = F.conv2d(y, w(2), b(2), padding=1)
        return F.conv2d(F.prelu(y, w(3)), w(4), b(4), padding=1)
    if kind == "head":     # up(bilinear)-conv-prelu-conv
        y = F.interpolate(x, scale_factor=2.0, mode="bilinear", align_corners=False)
        y = F.conv2d(y, w(1), b(1), padding=1)
        return F.conv2d(F.prelu(y, w(2)), w(3), b(3), padding=1)
    raise ValueError(kind)


def network_forward(sd, x1: torch.Tensor, x2: torch.Tensor, debug: dict = None) -> torch.Tensor:
    """Network.forward, sepconv_enhanced.py:605-706.  x1, x2: [B,3,H,W] fp32."""
    with torch.no_grad():
        W_, H_ = x1.shape[3], x1.shape[2]
        padr, padb = (2 - (W_ % 2)) % 2, (2 - (H_ % 2)) % 2
        one = F.pad(x1, [0, padr, 0, padb], mode="replicate")
        two = F.pad(x2, [0, padr, 0, padb], mode="replicate")
        stack = torch.stack([one, two], 1)
        mean = stack.view(stack.shape[0], -1).mean(1, True).view(-1, 1, 1, 1)
        std = stack.view(stack.shape[0], -1).std(1, True).view(-1, 1, 1, 1)
        seq = [(f - mean) / (std + 0.0000001) for f in (one, two)]
        feat = [F.conv2d(s, sd["netInput.weight"], sd["netInput.bias"], padding=1) for s in seq]
        lv = [torch.cat(feat, 1)]
        # Encode (:345-369): rows 1..4 = Ver(previous row); row 0 passes through
        for r in range(1, 5):
            lv.append(_basic(sd, f"netEncode.0.netVer.{r}", lv[r - 1], "enc"))
        # Decode (:447-497): Hor on rows 4..1, then Ver added on rows 3..1 (cropped by one where the coarser row is odd)
        for k, r in enumerate((4, 3, 2, 1)):
            lv[r] = _basic(sd, f"netDecode.0.netHor.{k}", lv[r], "hor")
        for k, r in enumerate((3, 2, 1), start=1):
            v = _basic(sd, f"netDecode.0.netVer.{k}", lv[r + 1], "dec")
            if v.shape[2] == lv[r].shape[2] + 1:
                v = v[:, :, :-1]
            if v.shape[3] == lv[r].shape[3] + 1:
                v = v[:, :, :, :-1]
            lv[r] = lv[r] + v
        out = lv[1]
        p = int(math.floor(0.5 * KSIZE))
        onep = F.pad(one, [p, p, p, p], mode="replicate")
        twop = F.pad(two, [p, p, p, p], mode="replicate")
        onep = torch.cat([onep, onep.new_ones([onep.shape[0], 1, onep.shape[2], onep.shape[3]])], 1)
        twop = torch.cat([twop, twop.new_ones([twop.shape[0], 1, twop.shape[2], twop.shape[3]])], 1)
        ver1 = _basic(sd, "netVerone", out, "head")
        ver2 = _basic(sd, "netVertwo", out, "head")
        hor1 = _basic(sd, "netHorone", out, "head")
        hor2 = _basic(sd, "netHortwo", out, "head")
        if debug is not None:
            debug.update(levels=lv, ver1=ver1, ver2=ver2, hor1=hor1, hor2=hor2, mean=mean, std=std)
        res = ops_ref.sepconv(onep, ver1, hor1) + ops_ref.sepconv(twop, ver2, hor2)
        norm = res[:, -1:, :, :].clone()
        norm[norm.abs() < 0.01] = 1.0
        res = res[:, :-1, :, :] / norm
        return res[:, :, :H_, :W_]
