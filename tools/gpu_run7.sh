#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; return $rc; }
run gpu_tests python -m pytest tests -q -m gpu -p no:cacheprovider -s
TAILN=60 run bench_layers python tools/bench_layers.py --batch 4 --json gpurun_out/layers_b4.json
TAILN=2 run bench python bench.py --steps 5 --warmup 3
run ncu_resconv ncu --set full --clock-control none --import-source on -k regex:tapconv_kernel -s 3 -c 1 -f -o gpurun_out/prof_resconv_b3 python tools/bench_layers.py --batch 4 --only 3:2 --iters 2
run ncu_warp ncu --set full --clock-control none --import-source on -k regex:warp_kernel -c 1 -f -o gpurun_out/prof_warp python bench.py --frames 3 --steps 1 --warmup 1 --no-cpu
run launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --frames 9 --steps 1 --warmup 1 --no-cpu
