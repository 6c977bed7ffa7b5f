#!/bin/bash
# e2e scaling probe: how the host-buffer path's time splits into a per-frame slope and a fixed head/tail
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=8 run gpu_tests python -m pytest tests -q -m gpu -p no:cacheprovider -k "not ref"
for f in 17 33 64 128; do
  TAILN=2 run bench_f$f python bench.py --steps 3 --warmup 2 --no-cpu --frames $f
  python - <<P
import json
for l in open("gpurun_out/bench_f$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("frames $f: dev ms", round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
P
done
