#!/bin/bash
# round 2 final pass (8 GPUs): configs [2] and [3] with the final defaults
mkdir -p gpurun_out; rm -f gpurun_out/r03y_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r03y_c4_n8.json 2> gpurun_out/r03y_c4_n8.err
timeout 600 $TR bench.py --gpus 8 --config bt8 --no-cpu-baseline > gpurun_out/r03y_bt8_n8.json 2> gpurun_out/r03y_bt8_n8.err
cut -c1-200 gpurun_out/r03y_c4_n8.json; cut -c1-200 gpurun_out/r03y_bt8_n8.json
