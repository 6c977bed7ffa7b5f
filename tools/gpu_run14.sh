#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/bin/mma_rate 2>&1 | tee gpurun_out/mma_rate2.log | head -12
