"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean us, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    a = agg.setdefault(d['Kernel Name'][:72], [0, 0.0, d['Grid Size'], d['Block Size']])
    a[0] += 1
    a[1] += v
tot = sum(t / n for n, t, _, _ in agg.values())
for k, (n, t, g, b) in agg.items():
    print(f"{n:4d} {t / n / 1e3:9.2f} us {100 * t / n / tot:5.1f}%  {g:>14s} {b:>12s} {k}")
print(f"sum of per-kernel means: {tot / 1e3:.1f} us")
