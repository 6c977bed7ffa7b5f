"""Host-side engine: owns a vfi_ctx, feeds it PyTorch device memory / host buffers through the C ABI.

PyTorch is plumbing here (device memory, streams); all arithmetic of the path runs in libvfi_b200.so.
"""
import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from ._lib import VfiError, check, lib

OPERAND = {"float32": 0, "float16": 0, "bfloat16": 1}
# names in IFNet("4.6").state_dict() order (reference rife_arch.py:177-218, :404-408)
BLOCK_C = (192, 128, 96, 64)


def state_dict_names(arch: str = "4.6"):
    names = []
    for b in range(5 if arch == "4.26" else 4):  # rife_arch.py:453-459: arch 4.26 has block0..block4
        p = f"block{b}."
        names += [p + "conv0.0.0.weight", p + "conv0.0.0.bias", p + "conv0.1.0.weight", p + "conv0.1.0.bias"]
        for j in range(8):
            q = p + f"convblock.{j}."
            names += [q + "beta", q + "conv.weight", q + "conv.bias"]
        names += [p + "lastconv.0.weight", p + "lastconv.0.bias"]
    if arch == "4.7":
        names += ["encode.0.weight", "encode.0.bias", "encode.1.weight", "encode.1.bias"]
    if arch in ("4.17", "4.26"):  # Head_417 - rife_arch.py:356-363; Head - rife_arch.py:378-385
        names += [f"encode.cnn{i}.{k}" for i in range(4) for k in ("weight", "bias")]
    return names


ARCH_CODE = {"4.6": 46, "4.7": 47, "4.17": 417, "4.26": 426}


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def host_copy_frames(frames: torch.Tensor, slots, out: torch.Tensor, threads: int = 0):
    """out[slots[i]] = frames[i, ..., :3] for every i with slots[i] >= 0 (CPU float32 tensors), on `threads` copy threads of
    the library (default VFI_COPY_THREADS or 12) - vfi_host_copy_frames."""
    assert not frames.is_cuda and not out.is_cuda and frames.dtype == torch.float32 and out.dtype == torch.float32
    assert frames.is_contiguous() and out.is_contiguous() and frames.dim() == 4 and tuple(out.shape[1:3]) == tuple(frames.shape[1:3])
    n, h, w, c = frames.shape
    sl = _i32(slots)
    assert len(sl) == n and (len(sl) == 0 or int(sl.max()) < out.shape[0]) and out.shape[3] == 3
    if threads <= 0:
        import os
        threads = int(os.environ.get("VFI_COPY_THREADS", "12"))
    check(lib().vfi_host_copy_frames(frames.data_ptr(), n, h, w, c, sl.ctypes.data, out.data_ptr(), int(threads)))


class Rife46Engine:
    """RIFE 4.6 / 4.7 / 4.17 / 4.26 (rife46.pth; rife47.pth, rife49.pth; rife417.pth; rife426.pth) on one B200.  `state_dict` maps the reference's parameter names to tensors (any float dtype)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int = 0, dtype: str = "float32", batch: int = 8,
                 arch: str = None):
        if not torch.cuda.is_available():
            raise VfiError("no CUDA device: this path has no CPU fallback")
        if dtype not in OPERAND:
            raise ValueError(f"dtype must be one of {list(OPERAND)}")
        self._L = lib()
        self.device = int(device)
        self._ctx = C.c_void_p()
        check(self._L.vfi_create(self.device, C.byref(self._ctx)))
        if arch is None:  # the 4.7 family (rife47.pth / rife49.pth) has the encode head, rife417.pth the Head_417 one
            arch = ("4.26" if "block4.conv0.0.0.weight" in state_dict else
                    "4.17" if "encode.cnn0.weight" in state_dict else
                    "4.7" if "encode.0.weight" in state_dict else "4.6")
        if arch not in ARCH_CODE:
            raise VfiError(f"RIFE arch {arch} is not built (4.6, 4.7, 4.17 and 4.26 are)")
        self.arch = arch
        names = state_dict_names(arch)
        missing = [n for n in names if n not in state_dict]
        if missing:
            raise KeyError(f"state_dict is not a RIFE {arch} checkpoint; missing {missing[:3]} ...")
        hold = [state_dict[n].detach().to("cpu", torch.float32).contiguous() for n in names]
        ptrs = (C.c_void_p * len(hold))(*[t.data_ptr() for t in hold])
        numel = (C.c_int64 * len(hold))(*[t.numel() for t in hold])
        check(self._L.vfi_rife_load(self._ctx, ARCH_CODE[arch], ptrs, numel, len(hold), OPERAND[dtype]))
        check(self._L.vfi_set_batch(self._ctx, int(batch)))
        self.dtype = dtype

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.vfi_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ device-resident path
    def forward(self, frames: torch.Tensor, f0: Sequence[int], f1: Sequence[int], t: Sequence[float],
                scale_factor: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: CUDA float32 [N,H,W,C>=3] NHWC; returns CUDA float32 [len(t),H,W,3] on the current stream."""
        assert frames.is_cuda and frames.dtype == torch.float32 and frames.is_contiguous() and frames.dim() == 4
        n, h, w, c = frames.shape
        f0a, f1a = _i32(f0), _i32(f1)
        ta = np.ascontiguousarray(np.asarray(t, dtype=np.float32))
        nt = len(ta)
        if out is None:
            out = torch.empty((nt, h, w, 3), dtype=torch.float32, device=frames.device)
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (nt, h, w, 3)
        st = torch.cuda.current_stream(frames.device).cuda_stream
        check(self._L.vfi_rife46_forward(self._ctx, frames.data_ptr(), n, h, w, c, f0a.ctypes.data, f1a.ctypes.data,
                                         ta.ctypes.data, nt, float(scale_factor), out.data_ptr(), st))
        return out

    # ------------------------------------------------------------------ host-buffer path (node / e2e)
    def interpolate_host(self, frames: torch.Tensor, f0, f1, t, out: torch.Tensor, out_slots=None,
                         frame_range=None, scale_factor: float = 1.0, frame_slots=None):
        """frames: CPU float32 [N,H,W,C]; out: CPU float32 [M,H,W,3]; task i -> out[out_slots[i]] (default i);
        frame_slots (optional, one per source frame, -1 = none): frame i is also copied unchanged to out[frame_slots[i]]."""
        assert not frames.is_cuda and frames.dtype == torch.float32 and frames.is_contiguous()
        assert not out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
        n, h, w, c = frames.shape
        assert tuple(out.shape[1:]) == (h, w, 3)
        f0a, f1a = _i32(f0), _i32(f1)
        ta = np.ascontiguousarray(np.asarray(t, dtype=np.float32))
        lo, hi = (0, n) if frame_range is None else frame_range
        slots = None
        if out_slots is not None:
            slots = _i32(out_slots)
            assert len(slots) == len(ta) and (len(slots) == 0 or int(slots.max()) < out.shape[0])
        else:
            assert out.shape[0] >= len(ta)
        fslots = None
        if frame_slots is not None:
            fslots = _i32(frame_slots)
            assert len(fslots) == n and int(fslots.max()) < out.shape[0]
        check(self._L.vfi_rife46_interpolate_host(
            self._ctx, frames.data_ptr(), n, h, w, c, int(lo), int(hi), f0a.ctypes.data, f1a.ctypes.data,
            ta.ctypes.data, None if slots is None else slots.ctypes.data, None if fslots is None else fslots.ctypes.data,
            len(ta), float(scale_factor), out.data_ptr()))
        return out

    # ------------------------------------------------------------------ primitives / hooks
    def warp(self, img: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
        """NHWC float32 CUDA: img [B,H,W,C], flow [B,H,W,2] (pixels) -> [B,H,W,C]  (rife_arch.warp semantics)."""
        assert img.is_cuda and flow.is_cuda and img.is_contiguous() and flow.is_contiguous()
        b, h, w, c = img.shape
        assert tuple(flow.shape) == (b, h, w, 2)
        out = torch.empty_like(img)
        st = torch.cuda.current_stream(img.device).cuda_stream
        check(self._L.vfi_warp_bilinear_border(self._ctx, img.data_ptr(), flow.data_ptr(), out.data_ptr(), b, h, w, c, st))
        return out

    def debug_layer(self, block, layer, x, out, out_mask=None, impl=0):
        b, h, w = x.shape[0], x.shape[1], x.shape[2]
        st = torch.cuda.current_stream(x.device).cuda_stream
        check(self._L.vfi_rife46_debug_layer(self._ctx, block, layer, x.data_ptr(), out.data_ptr(),
                                             None if out_mask is None else out_mask.data_ptr(), b, h, w, impl, st))
        return out

    def layer_plan(self, block, layer):
        st, nc, ns, sm, mc = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        check(self._L.vfi_rife46_layer_plan(self._ctx, block, layer, C.byref(st), C.byref(nc), C.byref(ns),
                                            C.byref(sm), C.byref(mc)))
        return dict(stages=st.value, n_cta=nc.value, nsplit=ns.value, smem_bytes=sm.value, macs_per_cell=mc.value)

    def debug_state(self, batch):
        """(flow [B,Hp,Wp,4], mask [B,Hp,Wp]) of the last internal pass as torch CUDA tensors."""
        hp, wp = C.c_int(), C.c_int()
        check(self._L.vfi_rife46_debug_state(self._ctx, None, None, batch, C.byref(hp), C.byref(wp)))
        flow = torch.empty((batch, hp.value, wp.value, 4), dtype=torch.float32, device=f"cuda:{self.device}")
        mask = torch.empty((batch, hp.value, wp.value), dtype=torch.float32, device=f"cuda:{self.device}")
        check(self._L.vfi_rife46_debug_state(self._ctx, flow.data_ptr(), mask.data_ptr(), batch, C.byref(hp),
                                             C.byref(wp)))
        return flow, mask

    def profile(self, enable: bool):
        """Start / stop per-kernel-group event timing of the forward schedule (vfi_rife_profile)."""
        check(self._L.vfi_rife_profile(self._ctx, 1 if enable else 0))

    def profile_read(self):
        """{group id: (total ms, spans)} since profile(True); ids: 10 * block + {0 front, 1 conv0.0, 2 conv0.1, 3 ResConv x8,
        4 lastconv}, 90 final."""
        ids = np.zeros(64, np.int32)
        ms = np.zeros(64, np.float32)
        cnt = np.zeros(64, np.int32)
        n = C.c_int()
        check(self._L.vfi_rife_profile_read(self._ctx, ids.ctypes.data, ms.ctypes.data, cnt.ctypes.data, 64, C.byref(n)))
        return {int(ids[i]): (float(ms[i]), int(cnt[i])) for i in range(n.value)}

    def set_batch(self, b):
        check(self._L.vfi_set_batch(self._ctx, int(b)))

    def launch_count(self) -> int:
        return int(self._L.vfi_launch_count(self._ctx))

    def sync(self):
        check(self._L.vfi_sync(self._ctx))


# ------------------------------------------------------------------------------------------------- FILM
def film_state_dict_names():
    """film_arch.Interpolator().state_dict() order (film_arch.py:91-100, :515-528, :550-565, :222-256, :391-393)."""
    names = []
    for i in range(4):
        for j in range(2):
            names += [f"extract.extract_sublevels.convs.{i}.{j}.0.weight", f"extract.extract_sublevels.convs.{i}.{j}.0.bias"]
    for prefix in ["predict_flow._predictor"] + [f"predict_flow._predictors.{k}" for k in range(3)]:
        for j in range(4):
            names += [f"{prefix}._convs.{j}.0.weight", f"{prefix}._convs.{j}.0.bias"]
        names += [f"{prefix}._convs.4.weight", f"{prefix}._convs.4.bias"]
    names += ["fuse.output_conv.weight", "fuse.output_conv.bias"]
    for k in range(4):
        names += [f"fuse.convs.{k}.0.weight", f"fuse.convs.{k}.0.bias"]
        for j in (1, 2):
            names += [f"fuse.convs.{k}.{j}.0.weight", f"fuse.convs.{k}.{j}.0.bias"]
    return names


class FilmEngine:
    """FILM (film_net_fp32.pt) on one B200: `state_dict` = the TorchScript module's / Interpolator's state_dict."""

    MAX_PAIRS = 16

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int = 0, dtype: str = "float32"):
        if not torch.cuda.is_available():
            raise VfiError("no CUDA device: this path has no CPU fallback")
        if dtype not in OPERAND:
            raise ValueError(f"dtype must be one of {list(OPERAND)}")
        self._L = lib()
        self.device = int(device)
        self._ctx = C.c_void_p()
        check(self._L.vfi_create(self.device, C.byref(self._ctx)))
        names = film_state_dict_names()
        missing = [n for n in names if n not in state_dict]
        if missing:
            raise KeyError(f"state_dict is not a FILM checkpoint; missing {missing[:3]} ...")
        hold = [state_dict[n].detach().to("cpu", torch.float32).contiguous() for n in names]
        ptrs = (C.c_void_p * len(hold))(*[t.data_ptr() for t in hold])
        numel = (C.c_int64 * len(hold))(*[t.numel() for t in hold])
        check(self._L.vfi_film_load(self._ctx, ptrs, numel, len(hold), OPERAND[dtype]))
        self.dtype = dtype

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.vfi_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, frames: torch.Tensor, f0: Sequence[int], f1: Sequence[int], clamp: bool = False,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: CUDA float32 [N,H,W,C>=3] NHWC -> CUDA float32 [len(f0),H,W,3]: the midpoint frame of every pair
        (Interpolator.forward, film_arch.py:458-459), optionally clamped to [0,1] as the node does (film/__init__.py:38)."""
        assert frames.is_cuda and frames.dtype == torch.float32 and frames.is_contiguous() and frames.dim() == 4
        n, h, w, c = frames.shape
        f0a, f1a = _i32(f0), _i32(f1)
        npairs = len(f0a)
        assert npairs == len(f1a) and 1 <= npairs <= self.MAX_PAIRS
        if out is None:
            out = torch.empty((npairs, h, w, 3), dtype=torch.float32, device=frames.device)
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (npairs, h, w, 3)
        st = torch.cuda.current_stream(frames.device).cuda_stream
        check(self._L.vfi_film_forward(self._ctx, frames.data_ptr(), n, h, w, c, f0a.ctypes.data, f1a.ctypes.data, npairs,
                                       1 if clamp else 0, out.data_ptr(), st))
        return out

    def midpoint(self, x0: torch.Tensor, x1: torch.Tensor) -> torch.Tensor:
        """[H,W,3] x 2 (CUDA float32) -> clamped midpoint [H,W,3]: one `model(x0, x1, dt).clamp(0, 1)` of the node."""
        pair = torch.stack([x0, x1]).contiguous()
        return self.forward(pair, [0], [1], clamp=True)[0]

    # ------------------------------------------------------------------ hooks
    def set_ref(self, use_ref: bool):
        check(self._L.vfi_film_debug_set_ref(self._ctx, 1 if use_ref else 0))

    def debug_conv(self, group, layer, src0, src1, out, B, H, W, impl=0, pitch0=None, pitch1=None, out_pitch=None):
        st = torch.cuda.current_stream(src0.device).cuda_stream
        check(self._L.vfi_film_debug_conv(
            self._ctx, group, layer, src0.data_ptr(), int(pitch0 if pitch0 is not None else src0.shape[-1]),
            None if src1 is None else src1.data_ptr(),
            0 if src1 is None else int(pitch1 if pitch1 is not None else src1.shape[-1]),
            out.data_ptr(), int(out_pitch if out_pitch is not None else out.shape[-1]), B, H, W, impl, st))
        return out

    def layer_plan(self, group, layer):
        v = [C.c_int() for _ in range(10)]
        check(self._L.vfi_film_layer_plan(self._ctx, group, layer, *[C.byref(x) for x in v]))
        keys = ("c0", "c1", "n_total", "ksize", "n_cta", "nsplit", "mt", "a_slots", "b_slots", "smem_bytes")
        return dict(zip(keys, [x.value for x in v]))

    def last_macs(self) -> int:
        return int(self._L.vfi_film_last_macs(self._ctx))

    def launch_count(self) -> int:
        return int(self._L.vfi_launch_count(self._ctx))

    def sync(self):
        check(self._L.vfi_sync(self._ctx))


# ------------------------------------------------------------------------------------------------- Sepconv
def sepconv_state_dict_names():
    """sepconv_enhanced.Network().state_dict() order (sepconv_enhanced.py:536-598)."""
    names = ["netInput.weight", "netInput.bias"]

    def basic(prefix, prelu0, c0, prelu1, c1):
        out = []
        if prelu0 is not None:
            out.append(f"{prefix}.netMain.{prelu0}.weight")
        out += [f"{prefix}.netMain.{c0}.weight", f"{prefix}.netMain.{c0}.bias", f"{prefix}.netMain.{prelu1}.weight",
                f"{prefix}.netMain.{c1}.weight", f"{prefix}.netMain.{c1}.bias"]
        return out

    for r in range(1, 5):
        names += basic(f"netEncode.0.netVer.{r}", 0, 1, 2, 3)
    for k in range(4):
        names += basic(f"netDecode.0.netHor.{k}", 0, 1, 2, 3)
    for k in range(1, 4):
        names += basic(f"netDecode.0.netVer.{k}", 0, 2, 3, 4)
    for head in ("netVerone", "netVertwo", "netHorone", "netHortwo"):
        names += basic(head, None, 1, 2, 3)
    return names


class SepconvEngine:
    """Sepconv (sepconv.pth) on one B200: `state_dict` = sepconv_enhanced.Network's."""

    MAX_PAIRS = 16

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int = 0, dtype: str = "float32"):
        if not torch.cuda.is_available():
            raise VfiError("no CUDA device: this path has no CPU fallback")
        if dtype not in OPERAND:
            raise ValueError(f"dtype must be one of {list(OPERAND)}")
        self._L = lib()
        self.device = int(device)
        self._ctx = C.c_void_p()
        check(self._L.vfi_create(self.device, C.byref(self._ctx)))
        names = sepconv_state_dict_names()
        missing = [n for n in names if n not in state_dict]
        if missing:
            raise KeyError(f"state_dict is not a Sepconv checkpoint; missing {missing[:3]} ...")
        hold = [state_dict[n].detach().to("cpu", torch.float32).contiguous() for n in names]
        ptrs = (C.c_void_p * len(hold))(*[t.data_ptr() for t in hold])
        numel = (C.c_int64 * len(hold))(*[t.numel() for t in hold])
        check(self._L.vfi_sepconv_load(self._ctx, ptrs, numel, len(hold), OPERAND[dtype]))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.vfi_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, frames: torch.Tensor, f0: Sequence[int], f1: Sequence[int],
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: CUDA float32 [N,H,W,C>=3] NHWC -> CUDA float32 [len(f0),H,W,3] = Network.forward(frame f0, frame f1)."""
        assert frames.is_cuda and frames.dtype == torch.float32 and frames.is_contiguous() and frames.dim() == 4
        n, h, w, c = frames.shape
        f0a, f1a = _i32(f0), _i32(f1)
        npairs = len(f0a)
        assert npairs == len(f1a) and 1 <= npairs <= self.MAX_PAIRS
        if out is None:
            out = torch.empty((npairs, h, w, 3), dtype=torch.float32, device=frames.device)
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (npairs, h, w, 3)
        st = torch.cuda.current_stream(frames.device).cuda_stream
        check(self._L.vfi_sepconv_forward(self._ctx, frames.data_ptr(), n, h, w, c, f0a.ctypes.data, f1a.ctypes.data, npairs,
                                          out.data_ptr(), st))
        return out

    def middle_frame(self, frame_0: torch.Tensor, frame_1: torch.Tensor) -> torch.Tensor:
        """NCHW [1,3,H,W] x 2 -> NCHW [1,3,H,W]: `model(frame_0, frame_1)` as generic_frame_loop calls it
        (sepconv/__init__.py:47-48)."""
        dev = torch.device("cuda", self.device)
        pair = torch.cat([frame_0, frame_1]).to(dev, torch.float32).permute(0, 2, 3, 1).contiguous()
        return self.forward(pair, [0], [1]).permute(0, 3, 1, 2)

    def set_ref(self, use_ref: bool):
        check(self._L.vfi_sepconv_debug_set_ref(self._ctx, 1 if use_ref else 0))

    def launch_count(self) -> int:
        return int(self._L.vfi_launch_count(self._ctx))

    def sync(self):
        check(self._L.vfi_sync(self._ctx))
