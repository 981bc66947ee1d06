#!/bin/bash
# round 2, pass e: templated conv epilogues, MCTS.playout, device rollout evaluator; NE=12 vs 16
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02e_pytest.log
timeout 120 python scripts/conv_microbench.py > gpurun_out/r02e_micro_ne12.log 2>&1
AZ_NN_NE=16 timeout 120 python scripts/conv_microbench.py > gpurun_out/r02e_micro_ne16.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02e_bench_c4.json 2> gpurun_out/r02e_bench_c4.err
AZ_NN_NE=16 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02e_bench_c4_ne16.json 2> gpurun_out/r02e_bench_c4_ne16.err
tail -3 gpurun_out/r02e_pytest.log
