"""ComfyUI node surface of the B200 Sepconv path: a drop-in for the reference's `Sepconv VFI` node
(vfi_models/sepconv/__init__.py:13-57): same attributes and kwargs; the model call inside
`generic_frame_loop(..., use_timestep=False)` is `SepconvEngine.middle_frame` (libvfi_b200.so), no fallback."""
import typing

import torch

from .engine import SepconvEngine
from .frame_loop import generic_frame_loop
from .node import InterpolationStateList, load_file_from_github_release

MODEL_TYPE = "sepconv"
CKPT_NAMES = ["sepconv.pth"]
_model_cache: typing.Dict[str, SepconvEngine] = {}


def _load_engine(ckpt_name: str) -> SepconvEngine:
    if ckpt_name not in _model_cache:
        model_path = load_file_from_github_release(MODEL_TYPE, ckpt_name)
        sd = torch.load(model_path, map_location="cpu", weights_only=False)   # sepconv/__init__.py:44
        _model_cache[ckpt_name] = SepconvEngine(sd, device=torch.cuda.current_device(), dtype="float32")
    return _model_cache[ckpt_name]


class SepconvVFI:
    @classmethod
    def INPUT_TYPES(s):
        return {
            "required": {
                "ckpt_name": (CKPT_NAMES,),
                "frames": ("IMAGE",),
                "clear_cache_after_n_frames": ("INT", {"default": 10, "min": 1, "max": 1000}),
                "multiplier": ("INT", {"default": 2, "min": 2, "max": 1000})
            },
            "optional": {
                "optional_interpolation_states": ("INTERPOLATION_STATES",)
            }
        }

    RETURN_TYPES = ("IMAGE",)
    FUNCTION = "vfi"
    CATEGORY = "ComfyUI-Frame-Interpolation/VFI"

    def vfi(
        self,
        ckpt_name: typing.AnyStr,
        frames: torch.Tensor,
        clear_cache_after_n_frames=10,
        multiplier: typing.SupportsInt = 2,
        optional_interpolation_states: InterpolationStateList = None,
        **kwargs
    ):
        engine = kwargs.pop("_engine", None) or _load_engine(ckpt_name)
        x = frames[..., :3].permute(0, 3, 1, 2)   # preprocess_frames, vfi_utils.py:139-140

        def return_middle_frame(frame_0, frame_1, timestep, model):
            return model.middle_frame(frame_0, frame_1)

        out = generic_frame_loop(type(self).__name__, x, clear_cache_after_n_frames, multiplier, return_middle_frame, engine,
                                 interpolation_states=optional_interpolation_states, use_timestep=False,
                                 dtype=torch.float32)
        return (out.permute(0, 2, 3, 1),)   # postprocess_frames, vfi_utils.py:142-143
