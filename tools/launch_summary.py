"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`):
   python tools/launch_summary.py X.csv > X_summary.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iV].replace(",", ""))
    us = v / 1e3 if r[iU] in ("ns", "nsecond") else (v * 1e3 if r[iU] in ("ms", "msecond") else v)
    name = r[iN].split("(")[0].replace("void ", "")
    tot[name] += us
    cnt[name] += 1
s = sum(tot.values())
print("kernel,launches,total_us,share")
for k, v in tot.most_common():
    print(f"{k},{cnt[k]},{v:.1f},{v / s:.4f}")
