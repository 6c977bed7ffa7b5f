#!/bin/bash
# v14: STG.256 + staged conv0.1 stores by default, half4 planes for the 4.7 / 4.17 fronts; window-stage count A/B;
# launch list and ncu captures of the dominant kernel
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-900; return $rc; }
TAILN=30 run gpu_tests_all python -m pytest tests -q -m gpu -p no:cacheprovider -s
TAILN=2 run bench python bench.py
VFI_STAGES_MAX=6 TAILN=1 run bench_stages6 python bench.py --no-cpu
VFI_STAGES_MAX=5 TAILN=1 run bench_stages5 python bench.py --no-cpu
VFI_STAGES_MAX=6 TAILN=6 run gpu_tests_stages6 python -m pytest tests/test_gpu_layers.py -q -m gpu -p no:cacheprovider
for a in 4.7 4.17 4.26; do TAILN=1 run bench_arch$a python bench.py --arch $a --no-cpu --steps 3 --warmup 3; done
TAILN=20 run layers python tools/bench_layers.py --batch 8 --json gpurun_out/r01_v14_layers_b8.json
VFI_STAGES_MAX=6 TAILN=20 run layers_stages6 python tools/bench_layers.py --batch 8 --json gpurun_out/r01_v14_layers_b8_stages6.json
TMO=300 TAILN=1 run launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_v14_launches.csv python bench.py --frames 9 --steps 1 --warmup 1 --no-cpu
NCU="ncu --set full --clock-control none --import-source on"
TAILN=2 run ncu_b3 $NCU -k regex:tapconv_kernel -s 3 -c 1 -o gpurun_out/r01_v14_resconv_b3 -f python tools/bench_layers.py --batch 8 --only 3:2 --iters 2
ls -la gpurun_out/*.ncu-rep
