#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ablate.py 2>&1 | tee gpurun_out/ablate.log | tail -60
