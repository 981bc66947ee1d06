#!/bin/bash
# round 2, pass g: full GPU suite (match harness, device replay), head with 2 CTAs/SM, train config on 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02g_pytest.log
timeout 120 python scripts/head_microbench.py 8192 8 8 > gpurun_out/r02g_head.log 2>&1
timeout 120 python scripts/head_microbench.py 1024 6 6 >> gpurun_out/r02g_head.log 2>&1
timeout 120 python scripts/head_microbench.py 16384 6 6 >> gpurun_out/r02g_head.log 2>&1
timeout 600 python bench.py --config bt8 --no-cpu-baseline > gpurun_out/r02g_bench_bt8.json 2> gpurun_out/r02g_bench_bt8.err
timeout 900 python bench.py --config train --generations 2 > gpurun_out/r02g_bench_train.json 2> gpurun_out/r02g_bench_train.err
timeout 900 python bench.py --config train --generations 2 --list-buffer > gpurun_out/r02g_bench_train_list.json 2> gpurun_out/r02g_bench_train_list.err
tail -5 gpurun_out/r02g_pytest.log; cat gpurun_out/r02g_head.log
