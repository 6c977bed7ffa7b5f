"""numpy restatement of the sampling rules the FILM kernels implement (csrc/film_elem.cu).  TEST INFRASTRUCTURE ONLY.

FILM's arithmetic is delegated to ATen like RIFE's; its call sites use three resampling operators with *size*
arguments on levels whose sizes are not related by exact factors of two (1080 -> 540 -> 270 -> 135 -> 67 -> 33 -> 16):
``F.interpolate(mode='bilinear', size=...)`` (film_arch.py:598, :611, :751), ``F.interpolate(mode='nearest',
size=...)`` (:290), ``F.avg_pool2d(2, 2)`` (:119, :674) and ``grid_sample`` behind ``warp`` (:677-724).  Each function
below is the index arithmetic of the corresponding kernel, line for line in float32 where the kernel uses float32;
``tests/test_film_primitives_np.py`` checks them against ATen, odd sizes included.
"""
import numpy as np


def bilinear_to_size(v: np.ndarray, H: int, W: int) -> np.ndarray:
    """flow_up_kernel: F.interpolate(v, size=(H, W), mode='bilinear') on NCHW (align_corners=False).
    scale = in / out in float32; src = max(scale * (dst + 0.5) - 0.5, 0); second tap clamped."""
    n, c, h, w = v.shape

    def taps(o, i):
        scale = np.float32(i) / np.float32(o)
        src = np.maximum(scale * (np.arange(o, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5), np.float32(0))
        i0 = np.minimum(src.astype(np.int64), i - 1)
        i1 = np.minimum(i0 + 1, i - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, np.float32(1) - l1, l1

    y0, y1, wy0, wy1 = taps(H, h)
    x0, x1, wx0, wx1 = taps(W, w)
    top = v[:, :, y0][:, :, :, x0] * wx0 + v[:, :, y0][:, :, :, x1] * wx1
    bot = v[:, :, y1][:, :, :, x0] * wx0 + v[:, :, y1][:, :, :, x1] * wx1
    return (top * wy0[None, None, :, None] + bot * wy1[None, None, :, None]).astype(np.float32)


def nearest_to_size(x: np.ndarray, H: int, W: int) -> np.ndarray:
    """nearest16_kernel: F.interpolate(x, size=(H, W), mode='nearest'): identity for equal sizes, dst >> 1 for an exact
    doubling, else min(floor(dst * float32(in / out)), in - 1)."""
    n, c, h, w = x.shape

    def idx(o, i):
        d = np.arange(o)
        if o == i:
            return d
        if o == 2 * i:
            return d >> 1
        scale = np.float32(i) / np.float32(o)
        return np.minimum(np.floor(d.astype(np.float32) * scale).astype(np.int64), i - 1)

    return x[:, :, idx(H, h)][:, :, :, idx(W, w)]


def avg_pool2(x: np.ndarray) -> np.ndarray:
    """pool_rgb_kernel / pool16_kernel: F.avg_pool2d(x, 2, 2): floor(H/2) x floor(W/2), a trailing odd row / column is
    dropped."""
    n, c, h, w = x.shape
    ho, wo = h // 2, w // 2
    x = x[:, :, :2 * ho, :2 * wo]
    return ((x[:, :, 0::2, 0::2] + x[:, :, 0::2, 1::2] + x[:, :, 1::2, 0::2] + x[:, :, 1::2, 1::2]) * np.float32(0.25))


def warp_pixel_offsets(img: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """warp16_kernel / film_taps: sample img (NCHW) at (x + flow[:,0], y + flow[:,1]), coordinates clamped to the image,
    bilinear; == film_arch.warp (:677-724), whose grid normalisation cancels against grid_sample's un-normalisation
    (align_corners=False, padding_mode='border')."""
    n, c, h, w = img.shape
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    out = np.empty_like(img)
    for b in range(n):
        sx = np.clip(xs + flow[b, 0], 0, w - 1).astype(np.float32)
        sy = np.clip(ys + flow[b, 1], 0, h - 1).astype(np.float32)
        x0 = np.floor(sx).astype(np.int64)
        y0 = np.floor(sy).astype(np.int64)
        x1 = np.minimum(x0 + 1, w - 1)
        y1 = np.minimum(y0 + 1, h - 1)
        wx = sx - x0
        wy = sy - y0
        im = img[b]
        out[b] = (im[:, y0, x0] * ((1 - wx) * (1 - wy)) + im[:, y0, x1] * (wx * (1 - wy)) +
                  im[:, y1, x0] * ((1 - wx) * wy) + im[:, y1, x1] * (wx * wy))
    return out
