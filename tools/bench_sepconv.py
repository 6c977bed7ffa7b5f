#!/usr/bin/env python
"""Sepconv harness (BASELINE.json configs[4]: Sepconv VFI 2x on a 4K clip; default here 1080p, --h 2160 --w 3840 for 4K).

    python tools/bench_sepconv.py [--pairs 2] [--steps 5] [--warmup 3] [--h 1080 --w 1920]

One step = one `vfi_sepconv_forward` over `--pairs` device-resident pairs (= that many interpolated frames), CUDA events
after >= 3 warm-up steps, inputs larger than L2.  Prints ONE JSON line: frames/s, launches per step, and the time of the
separable-convolution op alone (the kernel behind vfi_sepconv on one frame's shapes) as `op_ms`.
Not part of bench.py's driver contract (bench.py measures the RIFE north-star metric).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--w", type=int, default=1920)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.load_package()
    from cfi_b200.engine import SepconvEngine
    from cfi_b200 import ops as OPS
    from oracle import film as OF      # synthetic clip generator
    from oracle import sepconv as OS   # synthetic weights (no checkpoint ships)

    eng = SepconvEngine(OS.synthetic_state_dict(0), device=0)
    clip = OF.synthetic_clip(a.pairs + 1, a.h, a.w, seed=1234).cuda().contiguous()
    f0 = list(range(a.pairs))
    f1 = [i + 1 for i in f0]
    out = torch.empty((a.pairs, a.h, a.w, 3), dtype=torch.float32, device="cuda")
    for _ in range(max(a.warmup, 3)):
        eng.forward(clip, f0, f1, out=out)
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.forward(clip, f0, f1, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    line = {"metric": f"interpolated frames/sec @{a.w}x{a.h} Sepconv (one Network.forward per frame)",
            "value": a.pairs / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"Sepconv forward, {a.pairs} pairs of {a.h}x{a.w} per step, device resident"},
            "gpu_launches": (eng.launch_count() - l0) // a.steps, "finite": bool(torch.isfinite(out).all())}
    # the op alone, on one frame's shapes (two of these run per interpolated frame)
    try:
        he, we = a.h + a.h % 2, a.w + a.w % 2
        x = torch.rand(1, 4, he + 50, we + 50, device="cuda")
        v = torch.rand(1, 51, he, we, device="cuda")
        hz = torch.rand(1, 51, he, we, device="cuda")
        for _ in range(2):
            OPS.sepconv_func.apply(x, v, hz)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            OPS.sepconv_func.apply(x, v, hz)
        e1.record()
        torch.cuda.synchronize()
        line["op_ms"] = e0.elapsed_time(e1) / 3
    except Exception as e:  # the op harness is a convenience
        line["op_ms"] = None
        line["op_error"] = repr(e)[:200]
    print(json.dumps(line))
    eng.close()


if __name__ == "__main__":
    main()
