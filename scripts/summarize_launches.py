import collections, csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    name = row["Kernel Name"][:90]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%%  n=%4d  avg=%9.1f us  %s" % (t / tot * 100, n, t / n, k))
print("total us", tot)
