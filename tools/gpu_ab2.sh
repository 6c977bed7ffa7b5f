#!/bin/bash
mkdir -p gpurun_out
P=comfyui-frame-interpolation_b200/libvfi_b200.so
: > gpurun_out/ab2.log
for v in v8 cur v8 cur; do
  cp tools/ab/libvfi_$v.so $P
  echo "=== $v" | tee -a gpurun_out/ab2.log
  for l in 3:0 3:2 2:2; do
    timeout -s KILL 120 python tools/bench_layers.py --batch 4 --only $l --iters 10 2>&1 | tail -1 | cut -c1-130 | tee -a gpurun_out/ab2.log
  done
done
cp tools/ab/libvfi_cur.so $P
timeout -s KILL 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider -k "not ref" 2>&1 | tail -3 | tee -a gpurun_out/ab2.log
