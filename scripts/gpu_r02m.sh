#!/bin/bash
# round 2, pass m: elect.sync single-lane sections + stem trim
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02m_pytest.log
timeout 120 python scripts/conv_microbench.py > gpurun_out/r02m_micro.log 2>&1
timeout 120 python scripts/stem_microbench.py >> gpurun_out/r02m_micro.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02m_c4.json 2> gpurun_out/r02m_c4.err
tail -2 gpurun_out/r02m_pytest.log; cat gpurun_out/r02m_micro.log
