"""Aggregate GCCTRACE lines (bench.py --trace 1): per distinct GEMM shape, count, mean time, TFLOP/s, share."""
import re, sys
from collections import OrderedDict
agg = OrderedDict()
for line in open(sys.argv[1]):
    if not line.startswith("GCCTRACE"):
        continue
    kind = line.split()[1]
    kv = dict(re.findall(r"(\w+)=([\-\d\.]+)", line))
    key = (kind,) + tuple(kv[k] for k in ("N", "H", "W", "C", "R", "OH", "OW", "k", "s", "mode", "BN", "tiles", "ksplit", "kb", "stats"))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += float(kv["us"])
    a[2] += float(kv["us"]) * float(kv["tflops"]) * 1e6  # flop
tot = sum(a[1] for a in agg.values())
totf = sum(a[2] for a in agg.values())
print("GEMM launches %d, total %.2f ms, %.2f TFLOP -> %.0f TFLOP/s average" % (sum(a[0] for a in agg.values()), tot / 1e3, totf / 1e12, totf / tot / 1e6))
print("%-5s %3s %4s %4s %5s %5s %4s %4s %2s %2s %4s %5s %6s %3s %5s %2s | %4s %9s %8s %6s" % (
    "kind", "N", "H", "W", "C", "R", "OH", "OW", "k", "s", "mode", "BN", "tiles", "ks", "kb", "st", "cnt", "avg_us", "TFLOP/s", "share"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-5s %3s %4s %4s %5s %5s %4s %4s %2s %2s %4s %5s %6s %3s %5s %2s | %4d %9.1f %8.0f %5.1f%%" % (
        key + (a[0], a[1] / a[0], a[2] / a[1] / 1e6, 100 * a[1] / tot)))
