"""Per-region share of executed instructions / stall samples from `ncu --page source --csv` of one
kernel: python tools/ncu_src_regions.py report.ncu-rep <kernel regex> [window]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
win = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
print(rows[0][1][:120])
hdr = rows[1]
si, ii, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
blk = []
for r in rows[2:]:
    if len(r) < 10 or r[0] == "Kernel Name":
        break
    blk.append(r)
tot = sum(int(r[ii]) for r in blk)
tots = sum(int(r[smp]) for r in blk) or 1
print(len(blk), "SASS instructions; executed", tot, "samples", tots)
for st in range(0, len(blk), win):
    w = blk[st:st + win]
    c = sum(int(r[ii]) for r in w)
    s = sum(int(r[smp]) for r in w)
    if c == 0 and s == 0:
        continue
    ops = set()
    for r in w:
        op = [x for x in r[si].split() if not x.startswith("@")]
        if op and op[0].split(".")[0] in ("ATOMS", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "VOTE", "POPC", "ATOMG", "RED", "LD", "ST", "BREV"):
            ops.add(op[0].split(".")[0])
    print(f"{st:5d} inst {100 * c / tot:5.1f}%  samples {100 * s / tots:5.1f}%  {' '.join(sorted(ops))}")
