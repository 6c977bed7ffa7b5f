#!/bin/bash
# HEAD check after the 4-accumulator tapconv, arch 4.17 / 4.26, programmatic dependent launch, the store variants and the
# tiled op kernels: all GPU tests, smoke, default bench, A/B runs, the other archs' throughput, per-layer and op timings
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-900; return $rc; }
TAILN=25 run gpu_tests_all python -m pytest tests -q -m gpu -p no:cacheprovider -s || VFI_PDL=0 TAILN=25 run gpu_tests_nopdl python -m pytest tests -q -m gpu -p no:cacheprovider -s
VFI_STORE=2 TAILN=8 run gpu_tests_store2 python -m pytest tests/test_gpu_layers.py tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider
VFI_STORE=1 TAILN=4 run gpu_tests_store1 python -m pytest tests/test_gpu_layers.py -q -m gpu -p no:cacheprovider
TAILN=3 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=2 run bench python bench.py
VFI_PDL=0 TAILN=1 run bench_nopdl python bench.py --no-cpu
VFI_STORE=1 TAILN=1 run bench_store1 python bench.py --no-cpu
VFI_STORE=2 TAILN=1 run bench_store2 python bench.py --no-cpu
VFI_ISSUERS=2 TAILN=1 run bench_issuers2 python bench.py --no-cpu
TAILN=1 run bench_batch16 python bench.py --no-cpu --batch 16
VFI_STORE=2 TAILN=1 run bench_batch16_store2 python bench.py --no-cpu --batch 16
for a in 4.7 4.17 4.26; do TAILN=1 run bench_arch$a python bench.py --arch $a --no-cpu --steps 3 --warmup 3; done
TAILN=45 run layers python tools/bench_layers.py --batch 8 --json gpurun_out/r01_v13_layers_b8.json
VFI_STORE=2 TAILN=45 run layers_store2 python tools/bench_layers.py --batch 8 --json gpurun_out/r01_v13_layers_b8_store2.json
TAILN=12 run ops python tools/bench_ops.py --json gpurun_out/r01_v13_ops.json
VFI_SEPCONV_PY=0 VFI_OPS_TILED=0 TAILN=12 run ops_v0 python tools/bench_ops.py --json gpurun_out/r01_v13_ops_v0.json
VFI_SEPCONV_PY=1 TAILN=2 run ops_py1 python tools/bench_ops.py
VFI_SEPCONV_PY=3 TAILN=2 run ops_py3 python tools/bench_ops.py
