#!/bin/bash
# r03c: step_cycle_budget (time-based tail cap of k_step) sweep on BASELINE configs[2]
mkdir -p gpurun_out; rm -f gpurun_out/r03c_*.json
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q -k "sim_cap or bit_exact" 2>&1 | tail -4 > gpurun_out/r03c_tests.log
cat gpurun_out/r03c_tests.log
run() {  # name, extra flags
  timeout 400 python bench.py --no-cpu-baseline --steps 12 --warmup 3 $2 > gpurun_out/r03c_$1.json 2> gpurun_out/r03c_$1.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r03c_$1.json") if l.startswith("{")][-1]
    t = d["tree_roofline"] if "tree_roofline" in d else d["roofline"]
    print("%-14s %.3fM sims/s  ms/round %.4f  k_step %.4f ms  sims/row %s  sm %s" % ("$1", d["value"]/1e6, d["ms_per_round_trip"], d.get("tree_roofline", {}).get("avg_launch_ms", -1), round(d["sims_per_eval_slot"], 4), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1 failed", e)
PY
}
run cap8 "--sim-cap 8"
run cap4 "--sim-cap 4"
run b30k "--sim-cap 0 --cycle-budget 30000"
run b40k "--sim-cap 0 --cycle-budget 40000"
run b50k "--sim-cap 0 --cycle-budget 50000"
run b60k "--sim-cap 0 --cycle-budget 60000"
run b80k "--sim-cap 0 --cycle-budget 80000"
run cap8b50k "--sim-cap 8 --cycle-budget 50000"
