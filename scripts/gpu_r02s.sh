#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainer.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02s_pytest.log
timeout 900 python bench.py --config train --generations 3 > gpurun_out/r02s_train.json 2> gpurun_out/r02s_train.err
tail -3 gpurun_out/r02s_pytest.log; cut -c1-1200 gpurun_out/r02s_train.json; tail -3 gpurun_out/r02s_train.err
