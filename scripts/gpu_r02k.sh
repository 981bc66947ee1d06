#!/bin/bash
# round 2, pass k (8 GPUs): scaling of configs [2], [3], [4]
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02k_c4_n8.json 2> gpurun_out/r02k_c4_n8.err
timeout 600 $TR bench.py --gpus 8 --config bt8 --no-cpu-baseline > gpurun_out/r02k_bt8_n8.json 2> gpurun_out/r02k_bt8_n8.err
timeout 600 $TR bench.py --gpus 8 --config train --generations 3 > gpurun_out/r02k_train_n8.json 2> gpurun_out/r02k_train_n8.err
tail -c 300 gpurun_out/r02k_c4_n8.json; tail -c 300 gpurun_out/r02k_bt8_n8.json; tail -c 300 gpurun_out/r02k_train_n8.json
