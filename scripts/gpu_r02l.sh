#!/bin/bash
# round 2, pass l: full GPU suite + smoke + the bench lines of every BASELINE config on 1 GPU + sim-cap sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02l_c4.json 2> gpurun_out/r02l_c4.err
timeout 600 python bench.py --config bt6 --no-cpu-baseline > gpurun_out/r02l_bt6.json 2> gpurun_out/r02l_bt6.err
timeout 600 python bench.py --config bt6 --no-cpu-baseline --virtual-loss 8 > gpurun_out/r02l_bt6_vl8.json 2> gpurun_out/r02l_bt6_vl8.err
timeout 600 python bench.py --config bt8 --no-cpu-baseline > gpurun_out/r02l_bt8.json 2> gpurun_out/r02l_bt8.err
timeout 900 python bench.py --config train --generations 3 > gpurun_out/r02l_train.json 2> gpurun_out/r02l_train.err
for c in 4 12 16; do timeout 600 python bench.py --no-cpu-baseline --sim-cap $c > gpurun_out/r02l_c4_cap$c.json 2> gpurun_out/r02l_c4_cap$c.err; done
tail -3 gpurun_out/r02l_pytest.log; tail -2 gpurun_out/r02l_smoke.log
