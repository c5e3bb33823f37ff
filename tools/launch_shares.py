"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
ki, mi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= mi:
        continue
    try:
        v = float(r[mi].replace(',', ''))
    except ValueError:
        continue
    name = r[ki].split('(')[0].replace('void ', '')[:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print(f'{sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms total (cold-cache, serialised: shares, not absolutes)')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:62s} n={v[0]:4d} {v[1] / 1e6:9.3f} ms {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] / 1e3:8.1f} us')
