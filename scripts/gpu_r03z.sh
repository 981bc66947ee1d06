#!/bin/bash
# round 2 final pass (1 GPU): the driver's tiers (GPU tests, smoke, default bench, reference arm) + one line per BASELINE config
mkdir -p gpurun_out; rm -f gpurun_out/r03z_*
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r03z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03z_smoke.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03z_c4.json 2> gpurun_out/r03z_c4.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03z_ref.json 2> gpurun_out/r03z_ref.err
timeout 600 python bench.py --config bt6 --no-cpu-baseline --steps 100 > gpurun_out/r03z_bt6.json 2> gpurun_out/r03z_bt6.err
timeout 600 python bench.py --config bt6 --no-cpu-baseline --virtual-loss 8 --steps 40 > gpurun_out/r03z_bt6_vl8.json 2> gpurun_out/r03z_bt6_vl8.err
timeout 600 python bench.py --config bt8 --no-cpu-baseline > gpurun_out/r03z_bt8.json 2> gpurun_out/r03z_bt8.err
timeout 900 python bench.py --config train --generations 3 > gpurun_out/r03z_train.json 2> gpurun_out/r03z_train.err
tail -3 gpurun_out/r03z_pytest.log; tail -2 gpurun_out/r03z_smoke.log
for f in c4 ref bt6 bt6_vl8 bt8 train; do echo "== $f"; cut -c1-330 gpurun_out/r03z_$f.json; done
