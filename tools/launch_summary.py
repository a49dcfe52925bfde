"""Aggregate an ncu launch list (gpu__time_duration.sum CSV) by kernel: python tools/launch_summary.py in.csv > out.md"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
tot = 0.0
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name)
    ns = float(r[14].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
print("| kernel | launches | total ms | share | avg us |")
print("|---|---|---|---|---|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k[:90], n, ns / 1e6, 100 * ns / tot, ns / n / 1e3))
print("\ntotal: %d launches, %.3f ms (ncu-serialised, cold cache: shares, not absolutes)" % (len(rows), tot / 1e6))
