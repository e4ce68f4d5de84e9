#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time, launches and share per kernel.
usage: launch_summary.py <launches.csv> [launches_per_step]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
agg, seq = collections.OrderedDict(), []
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0].replace('void ', '').replace('b200::', '')
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    seq.append((name, v, r[gi]))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(seq)} launches, {tot:.1f} us total (per-launch times are cold-cache and serialised under ncu: compare SHARES)")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[1]:10.1f} us {a[0]:5d} launches {a[1] / tot * 100:6.1f}%  avg {a[1] / a[0]:8.2f} us  {k[:100]}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print("first launches in order:")
for s in seq[:n]:
    print(f"   {s[1]:8.2f} us  grid {s[2]:14s} {s[0][:90]}")
