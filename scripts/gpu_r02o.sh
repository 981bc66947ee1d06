#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do echo "AZ_NN_BLOCK_DEBUG=$d" >> gpurun_out/r02o_micro.log; AZ_NN_BLOCK_DEBUG=$d timeout 120 python scripts/block_microbench.py >> gpurun_out/r02o_micro.log 2>&1; done
cat gpurun_out/r02o_micro.log
