#!/bin/bash
# v15 round-end style measurement: all GPU tests, smoke, default bench (with cpu_baseline), launch list of one pass
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-900; return $rc; }
TAILN=30 run gpu_tests_all python -m pytest tests -q -m gpu -p no:cacheprovider -s
TAILN=3 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=2 run bench python bench.py
TAILN=2 run bench_bf16 python bench.py --no-cpu --dtype bfloat16
TMO=300 TAILN=1 run launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_v15_launches.csv python bench.py --frames 9 --steps 1 --warmup 1 --no-cpu
