#!/usr/bin/env python
"""Pipeline ablation of tapconv (diagnostic build, -DVFI_ABLATE): time one layer with parts of the pipeline off.
   masks: 1 no epilogue global stores/residual loads, 2 epilogue = tcgen05.ld + hand-shake, 4 no tcgen05.ld,
          8 no TMA, 16 no tcgen05.mma.   Build the library first (see tools/ablate.sh)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import rife46 as O  # noqa: E402

ge.load_package()
from cfi_b200 import _lib  # noqa: E402
_lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "libvfi_b200_ablate.so")
from cfi_b200.engine import Rife46Engine  # noqa: E402

BLOCK_C = (192, 128, 96, 64)
eng = Rife46Engine(O.synthetic_state_dict(0), 0, "float16")
print(_lib.lib().vfi_version().decode())
B, Hp, Wp = 8, 1088, 1920
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for blk, layer in ((3, 2), (3, 1), (2, 2)):
    s = (8, 4, 2, 1)[blk]
    c = BLOCK_C[blk]
    Hs, Ws = Hp // s, Wp // s
    ishape = {0: (B, Hs // 2, Ws // 2, 64), 1: (B, Hs // 4, Ws // 4, 2 * c), 2: (B, Hs // 4, Ws // 4, c)}[layer]
    oshape = {0: (B, Hs // 4, Ws // 4, 2 * c), 1: (B, Hs // 4, Ws // 4, c), 2: (B, Hs // 4, Ws // 4, c)}[layer]
    x = (0.1 * torch.randn(ishape, device="cuda")).half()
    out = torch.empty(oshape, dtype=torch.float16, device="cuda")
    om = torch.zeros(16, device="cuda")
    tiles = B * ((ishape[1] + 15) // 16) * ((ishape[2] + 7) // 8)
    pl = eng.layer_plan(blk, layer)
    for m in (0, 12, 24, 28, 1024 + 96):
        os.environ["VFI_ABLATE"] = str(m)
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            eng.debug_layer(blk, layer, x, out, om if False else None)
            e1.record()
            torch.cuda.synchronize()
            if it:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        items = tiles * pl["nsplit"]
        waves = items / 148.0
        print(f"b{blk} l{layer} n_cta {pl['n_cta']} x{pl['nsplit']} ablate={m:2d}: {ms * 1e3:7.1f} us  "
              f"= {ms * 1e-3 * 1.965e9 / waves:7.0f} cycles per tile-wave ({waves:.1f} waves)")
os.environ["VFI_ABLATE"] = "0"
eng.close()
