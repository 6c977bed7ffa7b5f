#!/bin/bash
# 2-GPU run: bench.py under torchrun (NCCL gather), reference arm under torchrun, N=1 for comparison
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-1500; }
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run bench_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3
run bench_n1 python bench.py --gpus 1 --steps 3 --warmup 3
run ref_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1
