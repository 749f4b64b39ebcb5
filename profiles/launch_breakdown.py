"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.  usage: launch_breakdown.py file.csv [launches_per_step]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
ik, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
seen, items = set(), []
for r in rows[h + 1:]:
    if len(r) <= iv or r[iid] in seen:
        continue
    seen.add(r[iid])
    items.append((re.sub(r'\(.*', '', r[ik]), float(r[iv].replace(',', ''))))
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in items:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"{len(items)} launches, {tot / 1e6:.2f} ms of kernel time (ncu per-launch durations: serialised, cold caches)")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v / 1e6:9.3f} ms {100 * v / tot:5.1f}%  n={n:4d}  avg {v / n / 1e3:8.1f} us  {k[:80]}")
