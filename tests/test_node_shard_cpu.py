"""RIFE_VFI.vfi's host logic on the CPU with stand-in engines: the task slices, frame ranges, output slots and pass-through
ownership it hands to the engines of several devices (node.set_devices) must tile the output exactly once - whatever the
skip list / multiplier list - and equal the single-device assembly.  (The engines' arithmetic is GPU-tested; here an engine
just records what it was asked to do.)"""
import numpy as np
import pytest
import torch


class FakeEngine:
    def __init__(self, dev, log):
        self.device, self.log = dev, log

    def interpolate_host(self, frames, f0, f1, t, out, out_slots=None, frame_range=None, scale_factor=1.0, frame_slots=None):
        lo, hi = frame_range
        assert all(lo <= a < hi and lo <= b < hi for a, b in zip(f0, f1)), "a task outside the engine's frame range"
        for a, b, tt, s in zip(f0, f1, t, out_slots):
            assert float(out[s, 0, 0, 0]) == -1.0, f"slot {s} written twice"
            out[s] = (1 - tt) * frames[a, ..., :3] + tt * frames[b, ..., :3]
        if frame_slots is not None:
            for i, s in enumerate(frame_slots):
                if s >= 0:
                    assert lo <= i < hi or True
                    assert float(out[s, 0, 0, 0]) == -1.0, f"pass-through slot {s} written twice"
                    out[s] = frames[i, ..., :3]
        self.log.append((self.device, len(t), frame_range))

    def close(self):
        pass


@pytest.mark.parametrize("devices", [[0], [0, 1], [0, 1, 2, 3]])
@pytest.mark.parametrize("multiplier,states", [(2, None), (3, ([1, 4], True)), ([2, 1, 4, 0, 3], None), (2, ([0, 2, 5], False))])
def test_node_sharding_tiles_the_output(pkg, monkeypatch, devices, multiplier, states):
    import cfi_b200.node as N
    log = []
    monkeypatch.setattr(N, "load_file_from_github_release", lambda model_type, ckpt_name: "unused")
    monkeypatch.setattr(N, "_engine_for", lambda ckpt, dtype, d, arch, path: FakeEngine(d, log))
    monkeypatch.setattr(N, "_alloc_output", lambda shape: torch.full(shape, -1.0))
    g = torch.Generator().manual_seed(3)
    fr = torch.rand(7, 6, 8, 4, generator=g)
    st = None if states is None else N.InterpolationStateList(list(states[0]), states[1])
    N.set_devices(devices)
    try:
        (out,) = N.RIFE_VFI().vfi("rife46.pth", fr, multiplier=multiplier, optional_interpolation_states=st)
    finally:
        N.set_devices(None)
    assert float(out.min()) >= 0.0, "an output slot was never written"
    # the single-device assembly of the same stand-in arithmetic
    tasks, mults = N.build_tasks(6, multiplier, st)
    exp, per = [], {}
    for p, t in tasks:
        per.setdefault(p, []).append(t)
    for p in range(6):
        exp.append(fr[p, ..., :3])
        for t in per.get(p, []):
            exp.append((1 - t) * fr[p, ..., :3] + t * fr[p + 1, ..., :3])
    exp.append(fr[6, ..., :3])
    assert torch.equal(out, torch.stack(exp))
    assert sum(n for _, n, _ in log) == len(tasks) and len(log) <= len(devices)
    if len(tasks) >= len(devices):  # contiguous, balanced slices
        sizes = [n for _, n, _ in log]
        assert max(sizes) - min(sizes) <= 1
