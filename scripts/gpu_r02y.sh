#!/bin/bash
# round 2, pass y: soak runs (arena overflow / stability over 20,000 round trips) + new geometry test + fused-block evaluator test
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "geometry or fused_blocks" 2>&1 | tail -4 > gpurun_out/r02y_pytest.log
timeout 900 python bench.py --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/r02y_soak_c4.json 2> gpurun_out/r02y_soak_c4.err
timeout 900 python bench.py --config bt8 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02y_soak_bt8.json 2> gpurun_out/r02y_soak_bt8.err
timeout 900 python bench.py --config bt6 --virtual-loss 8 --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/r02y_soak_bt6vl.json 2> gpurun_out/r02y_soak_bt6vl.err
python - <<'PY'
import json
for f in ["c4","bt8","bt6vl"]:
    for line in open("gpurun_out/r02y_soak_%s.json" % f):
        if line.startswith("{"):
            l=json.loads(line); print(f, "%.2fM sims/s" % (l["value"]/1e6), "ms/round %.4f" % l["ms_per_round_trip"], "games/s %.1f" % l["games_per_sec"], "overflow", l["overflow"], "peak nodes", l["peak_nodes_per_tree"], "cap", l["node_capacity"], l["clocks"]["sm_mhz"])
PY
tail -2 gpurun_out/r02y_pytest.log
