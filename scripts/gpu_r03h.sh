#!/bin/bash
# r03h: ncu launch list of BASELINE configs[1] (Breakthrough 6x6, 1,024 trees, exact mode): per-kernel GPU durations
mkdir -p gpurun_out; rm -f gpurun_out/r03h_*
timeout 300 python bench.py --config bt6 --steps 2 --warmup 3 --no-settle --no-graph --no-cpu-baseline > gpurun_out/r03h_plain.json 2> gpurun_out/r03h_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 260 --csv --log-file gpurun_out/r03h_bt6_launches.csv \
  python bench.py --config bt6 --steps 2 --warmup 3 --no-settle --no-graph --no-cpu-baseline > gpurun_out/r03h_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r03h_bt6_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%5.1f %%  %4d  %7.1f us  %s" % (100 * a[1] / tot, a[0], a[1] / a[0], k))
print("total per round trip: %.1f us" % (tot / 20))
PY
