#!/bin/bash
# r03s: soak of the final r02 build (arena overflow / stability): 20,000 Connect Four round trips, 10,000 of the Breakthrough
# configs (exact and virtual-loss), all with the default cycle budget
mkdir -p gpurun_out; rm -f gpurun_out/r03s_*
timeout 900 python bench.py --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/r03s_c4.json 2> gpurun_out/r03s_c4.err
timeout 900 python bench.py --config bt8 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r03s_bt8.json 2> gpurun_out/r03s_bt8.err
timeout 900 python bench.py --config bt6 --steps 2000 --warmup 5 --no-cpu-baseline > gpurun_out/r03s_bt6.json 2> gpurun_out/r03s_bt6.err
timeout 900 python bench.py --config bt6 --virtual-loss 8 --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/r03s_bt6vl.json 2> gpurun_out/r03s_bt6vl.err
python - <<'PY'
import json
for f in ["c4","bt8","bt6","bt6vl"]:
    try:
        l=[json.loads(x) for x in open("gpurun_out/r03s_%s.json" % f) if x.startswith("{")][-1]
        print(f, "%.2fM sims/s" % (l["value"]/1e6), "ms/round %.4f" % l["ms_per_round_trip"], "games/s %.1f" % l["games_per_sec"], "overflow", l["overflow"], "peak nodes", l["peak_nodes_per_tree"], "cap", l["node_capacity"], l["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/r03s_%s.err" % f).read()[-600:])
PY
