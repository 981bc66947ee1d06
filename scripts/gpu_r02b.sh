#!/bin/bash
# round 2, second GPU pass: the tcgen05 FC head for large action spaces
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02b_pytest_nn.log
timeout 300 python scripts/nn_error_probe.py > gpurun_out/r02b_nn_err.json 2> gpurun_out/r02b_nn_err.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1
timeout 600 python bench.py --config bt6 --no-cpu-baseline > gpurun_out/r02b_bench_bt6.json 2> gpurun_out/r02b_bench_bt6.err
timeout 600 python bench.py --config bt8 --no-cpu-baseline > gpurun_out/r02b_bench_bt8.json 2> gpurun_out/r02b_bench_bt8.err
tail -3 gpurun_out/r02b_pytest_nn.log; tail -2 gpurun_out/r02b_smoke.log
