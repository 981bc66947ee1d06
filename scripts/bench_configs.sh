# The BASELINE.json configurations on one GPU (run under gpurun): prints value / ms per step / clocks per config.
show() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M sims/s', round(d['ms_per_step'],4), 'ms/step  e2e', round(d['e2e']['value']/1e6,3), 'games/s', round(d['games_per_sec'],1), 'peak nodes', d['peak_nodes_per_tree'], 'overflow', d['overflow'], d['clocks'])"; }
python bench.py --no-cpu-baseline 2>/dev/null | show "c4_800_16384"
python bench.py --no-cpu-baseline --game "breakthrough(rows=6,columns=6)" --playouts 200 --trees 1024 --steps 2000 --warmup 2000 2>/dev/null | show "bt6_200_1024"
python bench.py --no-cpu-baseline --game breakthrough --playouts 800 --trees 4096 --steps 600 --warmup 1500 2>/dev/null | show "bt8_800_4096"
python bench.py --no-cpu-baseline --game breakthrough --playouts 800 --trees 16384 --steps 300 --warmup 1500 2>/dev/null | show "bt8_800_16384"
