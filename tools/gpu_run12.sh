#!/bin/bash
# MMA issue-rate micro-benchmark + ncu captures of the full-resolution kernels and of a ring layer
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=24 run mma_rate tools/bin/mma_rate
NCU="ncu --set full --clock-control none --import-source on"
TAILN=2 run ncu_front $NCU -k regex:"front2_kernel|front_kernel|final_kernel" -s 5 -c 5 -o gpurun_out/r01_v11_elementwise -f python bench.py --frames 9 --steps 1 --warmup 1 --no-cpu
TAILN=2 run ncu_ring $NCU -k regex:tapconv_kernel -s 3 -c 1 -o gpurun_out/r01_v11_resconv_b0_ring -f python tools/bench_layers.py --batch 8 --only 0:2 --iters 2
TAILN=2 run ncu_b3 $NCU -k regex:tapconv_kernel -s 3 -c 1 -o gpurun_out/r01_v11_resconv_b3 -f python tools/bench_layers.py --batch 8 --only 3:2 --iters 2
ls -la gpurun_out/*.ncu-rep
