"""Times the device SPSS construction (kmsc_spss_build, csrc/spss.cu) on the k-mer set of a random genome
and its mutated copy: build only (text stays on the device), build + fetch, string statistics per number of
matching rounds."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
import kmsc, synth
G = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
K, N, kb = 23, 14, 4
seqs = synth.phylogeny_sequences(2, G, p=0.002, seed=5)
km = np.unique(np.concatenate([synth.kmers_of(s, K, True) for s in seqs]))
ctx = kmsc.Context(0)
offs, keys = synth.csr_of(km, K, N, kb)
s = ctx.set_from_csr(K, N, kb, offs, keys)
print(f"{len(km)} canonical {K}-mers")
for rounds in (1, 2, 4, 8, 16):
    ctx.spss_build(s, True, rounds, fetch=False)
    t = time.time(); ns, nc = ctx.spss_build(s, True, rounds, fetch=False); ctx.sync(); dt = time.time() - t
    print(f"rounds {rounds:2d}: {ns} strings, {nc} chars ({nc/len(km):.3f} per k-mer), build {dt*1e3:.2f} ms = {len(km)/dt/1e6:.0f} M k-mers/s")
t = time.time(); strs = ctx.spss_build(s, True, 0); dt = time.time() - t
print(f"build + fetch + python strings: {dt*1e3:.1f} ms, longest string {max(map(len, strs))}")
