#!/bin/bash
# scripts/ab_bench2.sh "ENV1=a ENV2=b" "ENV1=c" ...  -> one bench line per environment string
for rep in 1 2; do
for envs in "$@"; do
  env $envs python bench.py --steps 400 --warmup 1000 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$envs', round(d['value']/1e6,3), round(d['ms_per_step'],4), {k: round(x*1e3,1) for k,x in d['roofline']['launch_ms_by_kind'].items()}, 'k_step', round(d['tree_roofline']['avg_launch_ms']*1e3,1))"
done; done
