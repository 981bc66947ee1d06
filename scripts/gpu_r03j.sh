#!/bin/bash
# r03j: per-tree cycle counts of k_step (az_debug_timing) with the cycle budget; Connect Four and the 1,024-tree Breakthrough config
mkdir -p gpurun_out
{ timeout 300 python scripts/kstep_tail.py connect_four 16384 800 16 64000
  timeout 300 python scripts/kstep_tail.py "breakthrough(rows=6,columns=6)" 1024 200 16 30000
  timeout 300 python scripts/kstep_tail.py breakthrough 8192 800 16 64000; } > gpurun_out/r03j_tail.log 2>&1
cat gpurun_out/r03j_tail.log
