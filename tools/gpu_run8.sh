#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=25 run gpu_tests python -m pytest tests -q -m gpu -p no:cacheprovider -s -k "not ref"
TAILN=2 run bench python bench.py --steps 3 --warmup 3 --no-cpu
