#!/bin/bash
# r03i: pool size of BASELINE configs[3] (8x8 Breakthrough; the config names no tree count)
mkdir -p gpurun_out; rm -f gpurun_out/r03i_*
run() {  # name, flags
  timeout 600 python bench.py --no-cpu-baseline --config bt8 --steps 20 --warmup 5 $2 > gpurun_out/r03i_$1.json 2> gpurun_out/r03i_$1.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r03i_$1.json") if l.startswith("{")][-1]
    print("%-10s %.3fM sims/s  ms/round %.4f  k_step %.4f ms  sims/row %.4f  sm %s %s" % ("$1", d["value"]/1e6, d["ms_per_round_trip"], d["tree_roofline"]["avg_launch_ms"], d["sims_per_eval_slot"], d["clocks"]["sm_mhz"], {k: round(v, 4) for k, v in d["roofline"]["launch_ms_by_kind"].items()}))
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r03i_$1.err").read()[-800:])
PY
}
run t8192 "--trees 8192"
run t16384 "--trees 16384"
run t12288 "--trees 12288"
