#!/bin/bash
# A/B two environment settings of the same build: scripts/ab_bench.sh VAR a b  -> value, ms/step, per-kernel times
var=$1; shift
for rep in 1 2; do
for v in "$@"; do
  env $EXTRA $var=$v python bench.py --steps 400 --warmup 1000 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$var=$v', round(d['value']/1e6,3), round(d['ms_per_step'],4), {k: round(x*1e3,1) for k,x in d['roofline']['launch_ms_by_kind'].items()}, 'k_step', round(d['tree_roofline']['avg_launch_ms']*1e3,1), 'nn', round(d['nn_roofline']['avg_forward_ms']*1e3,1))"
done; done
