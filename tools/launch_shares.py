#!/usr/bin/env python
"""Per-kernel time shares of ONE call from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    if unit in ('usecond', 'us'): v *= 1e3
    elif unit in ('msecond', 'ms'): v *= 1e6
    rows.append((row['Kernel Name'].split('(')[0], v, row['Grid Size']))
names = [r[0] for r in rows]
starts = [i for i, n in enumerate(names) if 'k_resize_roi_swap' in n]
ends = [i for i, n in enumerate(names) if 'k_post' in n]
s0 = starts[1]; e0 = [e for e in ends if e > s0][0]
call = rows[s0:e0 + 1]
tot = sum(v for _, v, _ in call)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, _ in call: agg[n][0] += 1; agg[n][1] += v
print(f'{len(call)} launches in one call; total {tot/1e3:.1f} us (serialised, cold cache)')
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{n:45s} x{c:3d} {v/1e3:9.1f} us {100*v/tot:5.1f}%')
if len(sys.argv) > 2:
    for i, (n, v, g) in enumerate(call): print(i, n[:36], f'{v/1e3:.1f}', g)
