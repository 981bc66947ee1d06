#!/bin/bash
# r03x (2 GPUs): bench output contract (one stdout line, also under torchrun where NCCL prints its banner) + the new contract test
mkdir -p gpurun_out; rm -f gpurun_out/r03x_*
timeout 900 python -m pytest tests/test_bench_contract.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r03x_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03x_c4_n2.json 2> gpurun_out/r03x_c4_n2.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 0 --cpu-seconds 2 > gpurun_out/r03x_ref_n2.json 2> gpurun_out/r03x_ref_n2.err
cat gpurun_out/r03x_pytest.log; wc -l gpurun_out/r03x_c4_n2.json gpurun_out/r03x_ref_n2.json; cut -c1-160 gpurun_out/r03x_c4_n2.json; grep -c "NCCL version" gpurun_out/r03x_c4_n2.err
