#!/bin/bash
# round 2, third GPU pass: conv N=160 + head finalize; numerics, benches of the three self-play configs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r02c_pytest_nn.log
timeout 300 python scripts/nn_error_probe.py > gpurun_out/r02c_nn_err.json 2> gpurun_out/r02c_nn_err.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench_c4.json 2> gpurun_out/r02c_bench_c4.err
timeout 600 python bench.py --config bt6 --no-cpu-baseline > gpurun_out/r02c_bench_bt6.json 2> gpurun_out/r02c_bench_bt6.err
timeout 600 python bench.py --config bt8 --no-cpu-baseline > gpurun_out/r02c_bench_bt8.json 2> gpurun_out/r02c_bench_bt8.err
for d in 1 2 4; do AZ_NN_DEBUG=$d timeout 120 python scripts/conv_microbench.py > gpurun_out/r02c_micro_$d.log 2>&1; done
timeout 120 python scripts/conv_microbench.py > gpurun_out/r02c_micro_0.log 2>&1
tail -3 gpurun_out/r02c_pytest_nn.log
