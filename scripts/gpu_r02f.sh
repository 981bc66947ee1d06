#!/bin/bash
# round 2, pass f: virtual-loss mode, head microbenchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "virtual_loss or external or rollout" 2>&1 | tail -30 > gpurun_out/r02f_pytest.log
for d in 0 1 2 4; do AZ_NN_HEAD_DEBUG=$d timeout 120 python scripts/head_microbench.py 8192 8 8 >> gpurun_out/r02f_head.log 2>&1; done
for d in 0 1 2 4; do AZ_NN_HEAD_DEBUG=$d timeout 120 python scripts/head_microbench.py 1024 6 6 >> gpurun_out/r02f_head.log 2>&1; done
timeout 120 python scripts/head_microbench.py 16384 6 6 >> gpurun_out/r02f_head.log 2>&1
timeout 120 python scripts/head_microbench.py 8192 6 6 >> gpurun_out/r02f_head.log 2>&1
for k in 0 2 4 8; do timeout 600 python bench.py --config bt6 --no-cpu-baseline --virtual-loss $k > gpurun_out/r02f_bench_bt6_vl$k.json 2> gpurun_out/r02f_bench_bt6_vl$k.err; done
tail -3 gpurun_out/r02f_pytest.log; cat gpurun_out/r02f_head.log
