#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; return $rc; }
run layers python -m pytest tests/test_gpu_layers.py tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "not ref"
run forward python -m pytest tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider
TAILN=60 run bench_layers python tools/bench_layers.py --batch 4 --json gpurun_out/layers_b4.json
TAILN=2 run bench python bench.py --steps 3 --warmup 3 --no-cpu
