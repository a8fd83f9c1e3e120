"""C5 shape: n sets of canonical 15-mers (<15,14,uint16>) -> dense-bitmap Gram (kmsc_bitmap_gram).
Wall clock around the synchronous C-ABI call (bitmap fill + Gram + D2H)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200")); sys.path.insert(0, str(ROOT / "tests"))
import kmsc, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
K, N, kb = 15, 14, 2
ctx = kmsc.Context(0)
seqs = synth.window_sequences(n, glen * 2, glen, p=0.005)
dev = []
for s in seqs:
    km = synth.kmer_set_of(s, K)
    offs, keys = synth.csr_of(km, K, N, kb)
    dev.append(ctx.set_from_csr(K, N, kb, offs, keys))
print(f"{n} sets, {dev[0].n_keys} keys each; bitmaps {n * 128 / 1024:.1f} GiB", flush=True)
for it in range(3):
    t = time.time(); W = ctx.bitmap_gram(dev); dt = time.time() - t
    print(f"iter {it}: {dt*1e3:.1f} ms  ({n * 2**27 / dt / 1e9:.0f} GB/s of bitmap bytes, {n*(n+1)/2 * 2**25 / dt / 1e12:.2f} T AND+POPC words/s)", flush=True)
ref = ctx.pair_counts(dev[:16])
print("parity with P3 on the first 16 sets:", np.array_equal(W[:16, :16], ref))
