"""numpy restatement of the ATen operators the reference's RIFE path calls.  TEST INFRASTRUCTURE ONLY.

The reference's arithmetic is delegated to PyTorch ATen (third-party, "torch" unpinned in
requirements-no-cupy.txt:1; this image: torch 2.11.0).  These are the published algorithms of
the five operators on the path, written as plain loops/array code so that the exact sampling
rules the CUDA kernels implement are stated independently of ATen.  ``tests/test_primitives_np.py``
checks each against ATen on small cases.  All layouts NCHW, fp32 (computed in fp64 then rounded).

call sites: rife_arch.py:64-70 (grid_sample), :238-266 (interpolate), :96-107 / :23 (conv2d),
:215-218 (conv_transpose2d, pixel_shuffle).
"""
import numpy as np


def bilinear_resize(x: np.ndarray, scale_factor: float) -> np.ndarray:
    """F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False), no antialias.

    Output size floor(in*s).  Source coordinate of output index d: ``src = (d + 0.5) / s - 0.5``,
    clamped below at 0; i0 = floor(src), i1 = min(i0 + 1, in - 1), lambda1 = src - i0.
    Consequences used by the kernels: for s = 1/k (k = 2, 4, 8 ...) the two taps are input
    pixels k*d + k/2 - 1 and k*d + k/2 with weights 1/2, 1/2; for s = 1 it is the identity.
    """
    n, c, h, w = x.shape
    oh, ow = int(np.floor(h * scale_factor)), int(np.floor(w * scale_factor))

    def taps(o, i):
        src = (np.arange(o, dtype=np.float64) + 0.5) / scale_factor - 0.5
        src = np.maximum(src, 0.0)
        i0 = np.minimum(np.floor(src).astype(np.int64), i - 1)
        i1 = np.minimum(i0 + 1, i - 1)
        l1 = src - i0
        return i0, i1, 1.0 - l1, l1

    y0, y1, wy0, wy1 = taps(oh, h)
    x0, x1, wx0, wx1 = taps(ow, w)
    xd = x.astype(np.float64)
    top = xd[:, :, y0][:, :, :, x0] * wx0 + xd[:, :, y0][:, :, :, x1] * wx1
    bot = xd[:, :, y1][:, :, :, x0] * wx0 + xd[:, :, y1][:, :, :, x1] * wx1
    return (top * wy0[None, None, :, None] + bot * wy1[None, None, :, None]).astype(np.float32)


def warp_border(img: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """rife_arch.warp == grid_sample(bilinear, padding_mode="border", align_corners=True) on a
    [-1,1] grid displaced by flow / ((size-1)/2): sample img at (x + fx, y + fy), the coordinate
    first clamped to [0, W-1] x [0, H-1], then 4-tap bilinear (the +1 tap of a coordinate sitting
    on the last row/column has weight 0)."""
    n, c, h, w = img.shape
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    out = np.empty((n, c, h, w), np.float32)
    for b in range(n):
        sx = np.clip(xs + flow[b, 0].astype(np.float64), 0, w - 1)
        sy = np.clip(ys + flow[b, 1].astype(np.float64), 0, h - 1)
        x0 = np.floor(sx).astype(np.int64)
        y0 = np.floor(sy).astype(np.int64)
        x1 = np.minimum(x0 + 1, w - 1)
        y1 = np.minimum(y0 + 1, h - 1)
        ax, ay = sx - x0, sy - y0
        im = img[b].astype(np.float64)
        v = (im[:, y0, x0] * (1 - ax) * (1 - ay) + im[:, y0, x1] * ax * (1 - ay)
             + im[:, y1, x0] * (1 - ax) * ay + im[:, y1, x1] * ax * ay)
        out[b] = v.astype(np.float32)
    return out


def conv2d_3x3(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int) -> np.ndarray:
    """Conv2d(k=3, stride, padding=1): out[o,y,x] = b[o] + sum_{c,ky,kx} w[o,c,ky,kx] * in[c, s*y+ky-1, s*x+kx-1]."""
    n, c, h, wd = x.shape
    oh, ow = (h + 2 - 3) // stride + 1, (wd + 2 - 3) // stride + 1
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    out = np.zeros((n, w.shape[0], oh, ow))
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, ky:ky + stride * oh:stride, kx:kx + stride * ow:stride]
            out += np.einsum("nchw,oc->nohw", patch, w[:, :, ky, kx].astype(np.float64))
    return (out + b[None, :, None, None]).astype(np.float32)


def conv_transpose2d_k4s2p1(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """ConvTranspose2d(k=4, stride=2, padding=1), weight [Cin, Cout, 4, 4]:
    out[o, 2i-1+ky, 2j-1+kx] += in[c,i,j] * w[c,o,ky,kx]  (output size 2H x 2W)."""
    n, c, h, wd = x.shape
    co = w.shape[1]
    full = np.zeros((n, co, 2 * h + 2, 2 * wd + 2))
    xd = x.astype(np.float64)
    for ky in range(4):
        for kx in range(4):
            full[:, :, ky:ky + 2 * h:2, kx:kx + 2 * wd:2] += np.einsum("nchw,co->nohw", xd, w[:, :, ky, kx].astype(np.float64))
    return (full[:, :, 1:-1, 1:-1] + b[None, :, None, None]).astype(np.float32)


def pixel_shuffle2(x: np.ndarray) -> np.ndarray:
    """PixelShuffle(2): out[c, 2h+i, 2w+j] = in[4c + 2i + j, h, w]."""
    n, c4, h, w = x.shape
    c = c4 // 4
    return x.reshape(n, c, 2, 2, h, w).transpose(0, 1, 4, 2, 5, 3).reshape(n, c, 2 * h, 2 * w)
