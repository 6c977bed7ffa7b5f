#!/bin/bash
# v15: residual loads without bank conflicts (explicit ld.shared, per-lane swizzled address), six stages for ResConv
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-900; return $rc; }
TAILN=6 run gpu_tests_lf python -m pytest tests/test_gpu_layers.py tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider
TAILN=1 run bench_v15 python bench.py --no-cpu
TAILN=20 run layers python tools/bench_layers.py --batch 8 --json gpurun_out/r01_v15_layers_b8.json
NCU="ncu --set full --clock-control none --import-source on"
TAILN=2 run ncu_b3 $NCU -k regex:tapconv_kernel -s 3 -c 1 -o gpurun_out/r01_v15_resconv_b3 -f python tools/bench_layers.py --batch 8 --only 3:2 --iters 2
