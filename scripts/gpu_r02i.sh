#!/bin/bash
# round 2, pass i (2 GPUs): the multi-rank bench path with the settle phase; train config with rank-0 training and with DDP
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_c4_n2.json 2> gpurun_out/r02i_bench_c4_n2.err
timeout 900 $TR bench.py --gpus 2 --config train --generations 2 > gpurun_out/r02i_train_n2.json 2> gpurun_out/r02i_train_n2.err
timeout 900 $TR bench.py --gpus 2 --config train --generations 2 --ddp > gpurun_out/r02i_train_n2_ddp.json 2> gpurun_out/r02i_train_n2_ddp.err
tail -c 600 gpurun_out/r02i_bench_c4_n2.json; tail -c 400 gpurun_out/r02i_train_n2.json; tail -c 400 gpurun_out/r02i_train_n2_ddp.json
