#!/bin/bash
# r03e: BT6 exact mode (1,024 trees), longer runs: count cap vs cycle budget
mkdir -p gpurun_out; rm -f gpurun_out/r03e_*.json
run() {  # name, flags
  timeout 400 python bench.py --no-cpu-baseline --config bt6 --steps 100 --warmup 5 $2 > gpurun_out/r03e_$1.json 2> gpurun_out/r03e_$1.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r03e_$1.json") if l.startswith("{")][-1]
    print("%-14s %.3fM sims/s  ms/round %.4f  k_step %.4f ms  sims/row %.4f  sm %s" % ("$1", d["value"]/1e6, d["ms_per_round_trip"], d["tree_roofline"]["avg_launch_ms"], d["sims_per_eval_slot"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1 failed", e)
PY
}
for rep in 1 2; do
run cap8_$rep "--sim-cap 8 --cycle-budget 0"
run b30k_$rep "--sim-cap 16 --cycle-budget 30000"
run b45k_$rep "--sim-cap 16 --cycle-budget 45000"
run b64k_$rep "--sim-cap 16 --cycle-budget 64000"
done
