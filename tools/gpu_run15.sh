#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=6 run gpu_tests python -m pytest tests -q -m gpu -p no:cacheprovider -k "not ref" || exit 1
TAILN=30 run layers_ring python tools/bench_layers.py --batch 8 --json gpurun_out/layers_b8_ring.json
TAILN=2 run bench python bench.py --steps 3 --warmup 3 --no-cpu
python - <<P
import json
for n in ("bench",):
  for l in open(f"gpurun_out/{n}.log"):
    if l.startswith("{"):
        d=json.loads(l); print(n, d["lib"], "dev ms", round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["clocks"])
P
timeout 300 python tools/ablate.py 2>&1 | grep -v "timed out" | tee gpurun_out/ablate.log | tail -30
