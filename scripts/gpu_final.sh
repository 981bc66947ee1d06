#!/bin/bash
# final verification of a round: the driver's GPU tiers (tests, smoke, default bench)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
tail -3 gpurun_out/final_pytest.log; tail -1 gpurun_out/final_smoke.log; cut -c1-400 gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_ref.json
