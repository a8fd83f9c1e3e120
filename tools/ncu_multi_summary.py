"""Per-kernel totals of an ncu --csv launch list taken with several --metrics (gpu__time_duration.sum,
smsp__inst_executed.sum, dram__bytes_read.sum, dram__bytes_write.sum).  usage: ncu_multi_summary.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
iK, iM, iV, iU = (hdr.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
B = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
for r in data:
    if len(r) <= iV:
        continue
    a, m, v = agg[r[iK][:80]], r[iM], float(r[iV].replace(',', ''))
    if m == 'gpu__time_duration.sum':
        a[0] += 1
        a[1] += v / 1e3 if r[iU] == 'ns' else (v if r[iU] == 'us' else v * 1e3)
    elif m == 'smsp__inst_executed.sum':
        a[2] += v
    elif m == 'dram__bytes_read.sum':
        a[3] += v * B[r[iU]]
    elif m == 'dram__bytes_write.sum':
        a[4] += v * B[r[iU]]
tot = sum(a[1] for a in agg.values())
print(f"total kernel time {tot:.0f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:10.0f} us {100 * a[1] / tot:5.1f}%  n={a[0]:4d} inst={a[2]:.3e} rd={a[3] / 1e9:7.2f}GB wr={a[4] / 1e9:7.2f}GB  {k}")
