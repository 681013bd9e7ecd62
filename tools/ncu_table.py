"""Turn `ncu --csv --metrics ...` logs into one row per kernel: launches, mean duration, DRAM bytes read / written.
usage: python tools/ncu_table.py log.csv [log2.csv ...]"""
import csv
import sys
from collections import defaultdict

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
rows = defaultdict(lambda: defaultdict(list))
for path in sys.argv[1:]:
    hdr = None
    for r in csv.reader(open(path, errors="replace")):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        name = r[hdr.index("Kernel Name")].split("(")[0]
        metric, unit, val = r[hdr.index("Metric Name")], r[hdr.index("Metric Unit")], r[hdr.index("Metric Value")]
        try:
            v = float(val.replace(",", "")) * UNIT.get(unit, 1)
        except ValueError:
            continue
        rows[name][metric].append(v)
print(f"{'kernel':44s} {'n':>4s} {'us':>9s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'GB/s':>8s} {'issue%':>7s} {'l2 hit%':>8s}")
for name, m in sorted(rows.items()):
    n = len(m.get("gpu__time_duration.sum", [])) or 1
    mean = lambda k: sum(m.get(k, [0])) / max(1, len(m.get(k, [0])))
    us, rd, wr = mean("gpu__time_duration.sum"), mean("dram__bytes_read.sum"), mean("dram__bytes_write.sum")
    print(f"{name[:44]:44s} {n:4d} {us:9.1f} {rd / 1e6:11.2f} {wr / 1e6:11.2f} {(rd + wr) / max(us, 1e-9) / 1e3:8.1f} "
          f"{mean('smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} {mean('lts__t_sector_hit_rate.pct'):8.1f}")
