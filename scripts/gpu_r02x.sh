#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "visit_counts_bit_exact or sim_cap" 2>&1 | tail -3 > gpurun_out/r02x_pytest.log
timeout 300 python scripts/overlap_probe2.py 8192 > gpurun_out/r02x_probe.log 2>&1
timeout 300 python scripts/overlap_probe2.py 16384 >> gpurun_out/r02x_probe.log 2>&1
AZ_NN_PDL=0 timeout 300 python scripts/overlap_probe2.py 8192 > gpurun_out/r02x_probe_nopdl.log 2>&1
tail -2 gpurun_out/r02x_pytest.log; cat gpurun_out/r02x_probe.log; echo "--- no PDL"; cat gpurun_out/r02x_probe_nopdl.log
