#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; return $rc; }
run layers_tc python -m pytest tests/test_gpu_layers.py -q -m gpu -k "tc or warp or plans" -p no:cacheprovider
run forward python -m pytest tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider
TAILN=60 run bench_layers python tools/bench_layers.py --batch 4 --json gpurun_out/layers_b4.json
TAILN=2 run bench python bench.py --steps 3 --warmup 3
best=""
for m in 0 1 2 3; do
  export VFI_TAPCONV_LAYOUT=1 VFI_TMA_MODE=$m
  if TMO=180 run tma_mode$m python -m pytest tests/test_gpu_layers.py -q -m gpu -k "tc" -p no:cacheprovider -x; then
    if [ -z "$best" ]; then best=$m; fi
  fi
done
echo "best TMA mode: '$best'" | tee gpurun_out/tma_best.txt
if [ -n "$best" ]; then
  export VFI_TAPCONV_LAYOUT=1 VFI_TMA_MODE=$best
  run tma_forward python -m pytest tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider
  TAILN=60 run tma_bench_layers python tools/bench_layers.py --batch 4 --json gpurun_out/layers_b4_tma.json
  TAILN=2 run tma_bench python bench.py --steps 3 --warmup 3 --no-cpu
  run tma_ncu ncu --set full --clock-control none --import-source on -k regex:tapconv_kernel -s 3 -c 1 -f -o gpurun_out/prof_resconv_b3_tma python tools/bench_layers.py --batch 4 --only 3:2 --iters 2
fi
unset VFI_TAPCONV_LAYOUT VFI_TMA_MODE
run ncu_resconv ncu --set full --clock-control none --import-source on -k regex:tapconv_kernel -s 3 -c 1 -f -o gpurun_out/prof_resconv_b3 python tools/bench_layers.py --batch 4 --only 3:2 --iters 2
