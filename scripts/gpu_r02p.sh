#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02p_micro.log
timeout 600 python -m pytest tests/test_gpu_nn.py -m gpu -q -x -k "residual_block" 2>&1 | tail -5 > gpurun_out/r02p_pytest.log
for d in 0 3; do echo "AZ_NN_BLOCK_DEBUG=$d" >> gpurun_out/r02p_micro.log; AZ_NN_BLOCK_DEBUG=$d timeout 120 python scripts/block_microbench.py >> gpurun_out/r02p_micro.log 2>&1; done
tail -3 gpurun_out/r02p_pytest.log; cat gpurun_out/r02p_micro.log
