"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
agg = collections.OrderedDict(); tot = 0.0
for r in rows[hi + 1:]:
    if len(r) <= iV: continue
    k = r[iK].split('(')[0][-70:]; v = float(r[iV].replace(',', '')); u = r[iU]
    v = v / 1e3 if u in ('ns', 'nsecond') else v * 1e3 if u in ('ms', 'msecond') else v
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"{sum(a[0] for a in agg.values())} launches, {tot/1e3:.3f} ms of kernel time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} n={n:5d} total {t/1e3:9.3f} ms ({100*t/tot:5.1f}%)  avg {t/n:9.1f} us")
