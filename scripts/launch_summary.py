#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of ONE forward into kernel / count / total us / share.
usage: python scripts/launch_summary.py gpurun_out/launches.csv [first_id last_id]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
head = rows[0]
iN, iV, iM, iD = head.index('Kernel Name'), head.index('Metric Value'), head.index('Metric Name'), head.index('ID')
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
agg = OrderedDict()
for r in rows[1:]:
    if r[iM] != 'gpu__time_duration.sum' or not (lo <= int(r[iD]) <= hi):
        continue
    name = r[iN].split('(')[0][:72]
    us = float(r[iV].replace(',', '')) / 1e3
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + us)
total = sum(t for _, t in agg.values())
for name, (n, t) in agg.items():
    print(f'{name:72s} x{n:2d} {t:8.1f} us {100 * t / total:5.1f}%')
print(f'TOTAL (serialised, cold cache) {total:.1f} us over {sum(n for n, _ in agg.values())} launches')
