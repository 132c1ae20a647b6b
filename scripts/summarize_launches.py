"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    grid = r.get("Grid Size", "")
    rows.append((name, ns, grid))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
# bench.py launches a cumsum (scan) marker right before the timed region: keep what follows the last marker
marks = [i for i, r in enumerate(rows) if "scan" in r[0].lower() or "cumsum" in r[0].lower()]
if marks:
    rows = rows[marks[-1] + 1:]
tot = sum(ns for _, ns, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for name, ns, _ in rows:
    agg[name][0] += 1
    agg[name][1] += ns
print("launches %d  total device time %.3f ms" % (len(rows), tot / 1e6))
print("%-70s %7s %10s %7s %9s" % ("kernel", "count", "ms", "share", "avg_us"))
for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %7d %10.3f %6.1f%% %9.1f" % (name[:70], cnt, ns / 1e6, 100 * ns / tot, ns / cnt / 1e3))
print("\ntop 25 single launches")
for name, ns, grid in sorted(rows, key=lambda r: -r[1])[:25]:
    print("%-60s %9.1f us  grid %s" % (name[:60], ns / 1e3, grid))
