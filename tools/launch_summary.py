"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean (us)."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("%-40s %7s %12s %10s %6s" % ("kernel", "count", "total_us", "mean_us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %7d %12.1f %10.2f %5.1f%%" % (k[:40], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
print("%-40s %7d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
