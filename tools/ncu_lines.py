"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source sass,cuda)
by CUDA source line: share of stall samples and of executed warp instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
lines = {}
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r
        iSamp = hdr.index('# Samples'); iInst = hdr.index('Instructions Executed')
        continue
    if hdr is None or len(r) <= iInst or r[0] == '':
        continue
    try:
        ln = int(r[0]); s = int(r[iSamp] or 0); n = int(r[iInst] or 0)
    except ValueError:
        continue
    a = lines.setdefault(ln, [r[1], 0, 0]); a[1] += s; a[2] += n
tot_s = sum(v[1] for v in lines.values()) or 1
tot_i = sum(v[2] for v in lines.values()) or 1
print('total samples', tot_s, 'warp instructions', tot_i)
for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:5d} samp {100*v[1]/tot_s:5.1f}% inst {100*v[2]/tot_i:5.1f}%  {v[0][:110]}")
