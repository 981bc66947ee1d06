#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py -m gpu -q -k "residual_block" 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -60 > gpurun_out/r02q_pytest.log
cat gpurun_out/r02q_pytest.log
