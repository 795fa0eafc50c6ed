"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python tools/launch_summary.py launches.csv "header line" > summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "").replace("scs::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iV].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("kernel | launches | total ms | share | avg ms")
for k, (n, ms) in agg.items():
    print("%s | %d | %.3f | %.1f%% | %.4f" % (k, n, ms, 100 * ms / tot, ms / n))
print("all | %d | %.3f | 100%% |" % (sum(v[0] for v in agg.values()), tot))
