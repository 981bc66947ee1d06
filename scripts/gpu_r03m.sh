#!/bin/bash
# r03m: more AZ_NN_DEBUG combinations on the plain conv (bit 1 = no input loads, 2 = no epilogue math / stores, 4 = no MMAs)
mkdir -p gpurun_out; rm -f gpurun_out/r03m_*
for d in 0 1 2 3 4 5 6 7; do AZ_NN_DEBUG=$d timeout 120 python scripts/conv_microbench.py 2>&1 | head -1 >> gpurun_out/r03m_micro.log; done
cat gpurun_out/r03m_micro.log
