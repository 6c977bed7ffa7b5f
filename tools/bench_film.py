#!/usr/bin/env python
"""FILM harness (BASELINE.json configs[2]: FILM 4x on a synthetic 1080p clip).

    python tools/bench_film.py [--pairs 2] [--steps 5] [--warmup 3] [--h 1080 --w 1920] [--layers] [--ref]

Every interpolated frame of FILM is one `Interpolator.forward` on a pair (the 4x schedule is 3 dependent calls per
source pair, pairs are independent), so the metric is forward calls per second = interpolated frames per second.
One step = one `vfi_film_forward` over `--pairs` device-resident pairs; timed with CUDA events after warm-up; inputs
larger than L2 at 1080p (2 x 25 MB per pair + ~6 GB of workspace traffic per call).  Prints ONE JSON line with
frames/s, the tensor-core TFLOP/s (unpadded MACs counted by the library) against the measured bf16 peak, and with
--layers the per-layer times of the streamconv launches (debug entry point, tcgen05 kernel).
--ref runs the same forward with every conv on the CUDA-core checker (a correctness aid, not a baseline).
Not part of bench.py's driver contract: bench.py measures the RIFE north-star metric.
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
    ap.add_argument("--layers", action="store_true")
    ap.add_argument("--ref", action="store_true")
    a = ap.parse_args()

    import __graft_entry__ as ge
    ge.load_package()
    from cfi_b200.engine import FilmEngine
    from oracle import film as OF   # weights and clip generators only (synthetic data, no checkpoint ships)

    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))

    eng = FilmEngine(OF.synthetic_state_dict(0), device=0, dtype="float32")
    eng.set_ref(a.ref)
    clip = OF.synthetic_clip(a.pairs + 1, a.h, a.w, seed=1234).cuda().contiguous()
    f0 = list(range(a.pairs))
    f1 = [i + 1 for i in f0]
    out = torch.empty((a.pairs, a.h, a.w, 3), dtype=torch.float32, device="cuda")
    for _ in range(max(a.warmup, 3)):
        eng.forward(clip, f0, f1, clamp=True, out=out)
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.forward(clip, f0, f1, clamp=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    macs = eng.last_macs()
    line = {
        "metric": "interpolated frames/sec @%dx%d FILM (one Interpolator.forward per frame)" % (a.w, a.h),
        "value": a.pairs / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": f"FILM forward, {a.pairs} pairs of {a.h}x{a.w} per step, device resident",
                   "impl": "checker (CUDA cores)" if a.ref else "tcgen05 streamconv"},
        "gpu_launches": eng.launch_count() - l0,
        "roofline": {"bound": "tensor", "achieved": 2 * macs / (ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": 2 * macs / (ms * 1e-3) / 1e12 / peak_tf, "traffic": None,
                     "note": "whole forward incl. the HBM-bound kernels; MACs of unpadded channels"},
        "gmacs_per_pair": macs / a.pairs / 1e9,
    }
    if a.layers:
        rows = []
        lvl = [(a.h >> l, a.w >> l) for l in range(7)]
        # (group, layer, level the layer runs on at its largest, label)
        todo = [(0, 2 * j + k, j, f"extract.{j}.{k}") for j in range(4) for k in (0, 1) if not (j == 0 and k == 0)]
        todo += [(1, 4 * p + c, p, f"flow.{'shared' if p == 3 else p}.{c}") for p in range(4) for c in range(4)]
        todo += [(2, 3 * k + c, 3 - k, f"fuse.{k}.{c}") for k in range(4) for c in range(3)]
        for g, layer, l, label in todo:
            plan = eng.layer_plan(g, layer)
            hh, ww = lvl[l]
            x0 = torch.zeros((1, hh, ww, plan["c0"]), dtype=torch.float16, device="cuda")
            x1 = torch.zeros((1, hh, ww, plan["c1"]), dtype=torch.float16, device="cuda") if plan["c1"] else None
            y = torch.empty((1, hh, ww, plan["n_total"]), dtype=torch.float16, device="cuda")
            for _ in range(3):
                eng.debug_conv(g, layer, x0, x1, y, 1, hh, ww)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                eng.debug_conv(g, layer, x0, x1, y, 1, hh, ww)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 5 * 1e3
            fl = 2.0 * plan["ksize"] ** 2 * (plan["c0"] + plan["c1"]) * plan["n_total"] * hh * ww
            rows.append(dict(layer=label, level=l, us=round(us, 1), tflops_padded=round(fl / us / 1e6, 1), **plan))
        line["layers"] = rows
    print(json.dumps(line))


if __name__ == "__main__":
    main()
