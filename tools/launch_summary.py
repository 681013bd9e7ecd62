"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        agg[d["Kernel Name"].split("(")[0][:60]][0] += 1
        agg[d["Kernel Name"].split("(")[0][:60]][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} {'n':>6s} {'total_us':>12s} {'avg_us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:60s} {v[0]:6d} {v[1] / 1e3:12.1f} {v[1] / 1e3 / v[0]:9.2f} {100 * v[1] / tot:5.1f}%")
