#!/bin/bash
# round-end style measurement: all GPU tests, smoke, default bench (with cpu_baseline), launch list, ncu of the dominant kernel
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; return $rc; }
TAILN=8 run gpu_tests_all python -m pytest tests -q -m gpu -p no:cacheprovider -s
TAILN=3 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=2 run bench python bench.py
TAILN=40 run layers python tools/bench_layers.py --batch 8 --json gpurun_out/r01_v12_layers_b8.json
TMO=300 TAILN=1 run launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_v12_launches.csv python bench.py --frames 9 --steps 1 --warmup 1 --no-cpu
NCU="ncu --set full --clock-control none --import-source on"
TAILN=2 run ncu_b3 $NCU -k regex:tapconv_kernel -s 3 -c 1 -o gpurun_out/r01_v12_resconv_b3 -f python tools/bench_layers.py --batch 8 --only 3:2 --iters 2
TAILN=2 run ncu_front $NCU -k regex:"front2_kernel|front_kernel|final_kernel" -s 5 -c 5 -o gpurun_out/r01_v12_elementwise -f python bench.py --frames 9 --steps 1 --warmup 1 --no-cpu
ls -la gpurun_out/*.ncu-rep
