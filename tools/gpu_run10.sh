#!/bin/bash
# parity + bench + launch list of one short step
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=30 run gpu_tests python -m pytest tests -q -m gpu -p no:cacheprovider -s -k "not ref"
TAILN=2 run bench python bench.py --steps 3 --warmup 3 --no-cpu
python - <<P
import json
for l in open("gpurun_out/bench.log"):
    if l.startswith("{"):
        d=json.loads(l); print("dev ms", round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
P
TMO=300 TAILN=1 run launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --frames 9 --steps 1 --warmup 1 --no-cpu
