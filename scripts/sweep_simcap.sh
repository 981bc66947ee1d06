for cap in 3 4 6 8 12; do python bench.py --steps 600 --warmup 1500 --no-cpu-baseline --sim-cap $cap 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cap $cap', round(d['value']/1e6,3), round(d['ms_per_step'],4), 'k_step', round(d['tree_roofline']['avg_launch_ms']*1e3,1), d['clocks']['sm_mhz'])"; done
