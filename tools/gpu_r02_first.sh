#!/bin/bash
# First GPU session of r02: everything r01 could not measure.  One gpurun call (about 12 min of box time):
#   gpurun --timeout 1500 -- bash tools/gpu_r02_first.sh
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-1200; return $rc; }
TAILN=30 run film_check_full python tools/film_gpu_check.py
TAILN=12 run sepconv_check python tools/sepconv_gpu_check.py
TAILN=30 TMO=900 run gpu_tests_all python -m pytest tests -q -m gpu -p no:cacheprovider
TAILN=2 run bench_film python tools/bench_film.py --frames 5 --steps 2 --layers
# the cluster-multicast variant of streamconv (written blind at the end of r01, default off): parity first, then the A/B
VFI_SC_CLUSTER=2 TAILN=20 run film_check_cluster2 python tools/film_gpu_check.py --quick
VFI_SC_CLUSTER=2 TAILN=2 run bench_film_cluster2 python tools/bench_film.py --frames 5 --steps 2 --no-cpu
TAILN=2 run bench_sepconv python tools/bench_sepconv.py --h 1080 --w 1920 --steps 3
TAILN=2 run bench python bench.py
TAILN=20 run bench_ops python tools/bench_ops.py --small --json gpurun_out/r02_bench_ops.json
TMO=300 TAILN=1 run film_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_film_launches.csv python tools/bench_film.py --frames 2 --multiplier 2 --pairs 1 --steps 1 --warmup 3 --no-cpu
# the widest streamconv layer of one 1080p forward (fusion level 3: cat(1984, 512) -> 512), full metric set
TMO=600 TAILN=1 run film_ncu ncu --set full --clock-control none --import-source on -k regex:streamconv_kernel -s 60 -c 1 -o gpurun_out/r02_film_streamconv python tools/bench_film.py --frames 2 --multiplier 2 --pairs 1 --steps 1 --warmup 3 --no-cpu
