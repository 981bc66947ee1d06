#!/bin/bash
# r03d: step_cycle_budget on the other configs (BT6 exact 1,024 trees, BT8 8,192 trees)
mkdir -p gpurun_out; rm -f gpurun_out/r03d_*.json
run() {  # name, flags
  timeout 400 python bench.py --no-cpu-baseline --steps 12 --warmup 3 $2 > gpurun_out/r03d_$1.json 2> gpurun_out/r03d_$1.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r03d_$1.json") if l.startswith("{")][-1]
    print("%-14s %.3fM sims/s  ms/round %.4f  k_step %.4f ms  sims/row %.4f  sm %s" % ("$1", d["value"]/1e6, d["ms_per_round_trip"], d["tree_roofline"]["avg_launch_ms"], d["sims_per_eval_slot"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1 failed", e)
PY
}
run bt6_cap8 "--config bt6 --sim-cap 8 --cycle-budget 0"
run bt6_b20k "--config bt6 --sim-cap 16 --cycle-budget 20000"
run bt6_b30k "--config bt6 --sim-cap 16 --cycle-budget 30000"
run bt6_b45k "--config bt6 --sim-cap 16 --cycle-budget 45000"
run bt6_b64k "--config bt6 --sim-cap 16 --cycle-budget 64000"
run bt8_cap8 "--config bt8 --sim-cap 8 --cycle-budget 0"
run bt8_b40k "--config bt8 --sim-cap 16 --cycle-budget 40000"
run bt8_b64k "--config bt8 --sim-cap 16 --cycle-budget 64000"
run c4_default ""
