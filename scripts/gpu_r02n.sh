#!/bin/bash
# round 2, pass n: fused residual block kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py -m gpu -q -x -k "residual_block" 2>&1 | tail -25 > gpurun_out/r02n_pytest.log
timeout 120 python scripts/block_microbench.py > gpurun_out/r02n_micro.log 2>&1
tail -5 gpurun_out/r02n_pytest.log; cat gpurun_out/r02n_micro.log
