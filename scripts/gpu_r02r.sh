#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02r_micro.log
for d in 3 7 15; do echo "AZ_NN_BLOCK_DEBUG=$d" >> gpurun_out/r02r_micro.log; AZ_NN_BLOCK_DEBUG=$d timeout 120 python scripts/block_microbench.py 2>&1 | grep fused >> gpurun_out/r02r_micro.log; done
cat gpurun_out/r02r_micro.log
