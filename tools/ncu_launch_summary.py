"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean
time per kernel name (torch's own kernels filtered unless ALL=1)."""
import collections, csv, os, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki]
    if not os.environ.get("ALL") and ("at::" in k or "elementwise" in k):
        continue
    a = agg.setdefault(k[:90], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(t for _, t in agg.values())
for k, (c, t) in agg.items():
    print(f"{c:5d} x {t / c / 1e3:9.1f} us = {t / 1e3:10.1f} us ({100 * t / tot:5.1f} %)  {k}")
print(f"total {tot / 1e3:.1f} us")
