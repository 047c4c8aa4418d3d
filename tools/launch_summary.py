"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list: per-kernel count / total / mean of the LAST invocation third."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u in ("nsecond", "ns") else v if u in ("usecond", "us") else v * 1e3 if u in ("msecond", "ms") else v
        rows.append((r["Kernel Name"].split("(")[0].split("<")[0], v))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
per = len(rows) // reps
last = rows[-per:]
agg = collections.OrderedDict()
for k, v in last:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in last)
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n={n:3d} {t:9.1f} us  {100*t/tot:5.1f}%")
print(f"{'TOTAL':40s} n={len(last):3d} {tot:9.1f} us")
