#!/bin/bash
mkdir -p gpurun_out
for na in 2 4; do for eo in 0 1; do
  echo "=== NACC=$na EPI_ORDER=$eo"
  for l in 3:0 3:1 3:2 2:2; do
    VFI_NACC=$na VFI_EPI_ORDER=$eo timeout -s KILL 120 python tools/bench_layers.py --batch 4 --only $l --iters 10 2>&1 | tail -1 | cut -c1-140
  done
done; done 2>&1 | tee gpurun_out/ab.log
