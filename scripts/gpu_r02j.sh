#!/bin/bash
# round 2, pass j: FC head rendezvous mode (logits stay in TMEM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02j_pytest.log
for m in 1 0; do
AZ_NN_HEAD_MODE=$m timeout 120 python scripts/head_microbench.py 8192 8 8 >> gpurun_out/r02j_head.log 2>&1
AZ_NN_HEAD_MODE=$m timeout 120 python scripts/head_microbench.py 16384 8 8 >> gpurun_out/r02j_head.log 2>&1
AZ_NN_HEAD_MODE=$m timeout 120 python scripts/head_microbench.py 1024 6 6 >> gpurun_out/r02j_head.log 2>&1
AZ_NN_HEAD_MODE=$m timeout 120 python scripts/head_microbench.py 16384 6 6 >> gpurun_out/r02j_head.log 2>&1
done
timeout 600 python bench.py --config bt8 --no-cpu-baseline > gpurun_out/r02j_bench_bt8.json 2> gpurun_out/r02j_bench_bt8.err
timeout 600 python bench.py --config bt6 --no-cpu-baseline --virtual-loss 8 > gpurun_out/r02j_bench_bt6_vl8.json 2> gpurun_out/r02j_bench_bt6_vl8.err
tail -3 gpurun_out/r02j_pytest.log; cat gpurun_out/r02j_head.log
