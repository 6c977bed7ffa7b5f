#!/bin/bash
# First-contact run on the GPU box: every stage in its own process with a timeout, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 15 gpurun_out/$name.log; }
run layers_ref python -m pytest tests/test_gpu_layers.py -q -m gpu -k "ref or warp or plans" -p no:cacheprovider
run layers_tc python -m pytest tests/test_gpu_layers.py -q -m gpu -k "tc" -p no:cacheprovider
run forward python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_layers python tools/bench_layers.py --batch 4 --json gpurun_out/layers_b4.json
