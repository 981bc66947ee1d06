#!/bin/bash
# round 2, first GPU pass: tests, evaluator numerics probe, steady-state bench lines of configs [2], [1], [3]
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02a_pytest.log
python scripts/nn_error_probe.py > gpurun_out/r02a_nn_err.json 2> gpurun_out/r02a_nn_err.err
python bench.py > gpurun_out/r02a_bench_c4.json 2> gpurun_out/r02a_bench_c4.err
python bench.py --config bt6 --no-cpu-baseline > gpurun_out/r02a_bench_bt6.json 2> gpurun_out/r02a_bench_bt6.err
python bench.py --config bt8 --no-cpu-baseline > gpurun_out/r02a_bench_bt8.json 2> gpurun_out/r02a_bench_bt8.err
tail -3 gpurun_out/r02a_pytest.log
