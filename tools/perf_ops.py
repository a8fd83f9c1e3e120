"""Wall-clock timing of the synchronous C-ABI calls around P3 (P2 decode, P4 split / union /
diff) on C2-sized sets (10 M canonical 23-mers each). Not the bench: a development probe."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200")); sys.path.insert(0, str(ROOT / "tests"))
import kmsc, synth

K, N, kb = 23, 14, 4
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = kmsc.Context(0)
seqs = synth.phylogeny_sequences(3, G + K - 1, p=0.002)
sets = []
for s in seqs:
    w = synth.pack_words(s) if hasattr(synth, "pack_words") else None
    strs_offs = np.array([0, len(s)], np.int64)
    t = time.time()
    d = ctx.set_from_spss(K, N, kb, [synth.to_ascii(s).decode()])
    print(f"set_from_spss (ASCII, {len(s)} bases): {1e3*(time.time()-t):.2f} ms, {d.n_keys} keys", flush=True)
    sets.append(d)

def timeit(name, fn, reps=10, bytes_=None):
    fn(); ts = []
    for _ in range(reps):
        t = time.time(); r = fn(); ts.append(time.time() - t)
        if isinstance(r, tuple):
            for x in r: x.free()
        elif hasattr(r, "free"): r.free()
    best, med = min(ts), sorted(ts)[len(ts)//2]
    extra = f"  {bytes_/best/1e9:.0f} GB/s algorithmic (best)" if bytes_ else ""
    print(f"{name}: best {1e3*best:.3f} ms, median {1e3*med:.3f} ms{extra}", flush=True)

a, b = sets[0], sets[1]
i, am, bm = ctx.pair_split(a, b)
B_s = (a.n_keys + b.n_keys + i.n_keys + am.n_keys + bm.n_keys) * kb + 5 * ((1 << N) + 1) * 4
print("split sizes", i.n_keys, am.n_keys, bm.n_keys, "B_s =", B_s)
for x in (i, am, bm): x.free()
timeit("pair_split (3 outputs)", lambda: ctx.pair_split(a, b), bytes_=B_s)
timeit("set_union (2 sets)", lambda: ctx.set_union([a, b]), bytes_=(a.n_keys + b.n_keys) * kb * 2)
timeit("set_diff", lambda: ctx.set_diff(a, b), bytes_=(a.n_keys + b.n_keys) * kb)
timeit("set_hash", lambda: a.Hash(), bytes_=a.n_keys * kb)
timeit("pair_counts (3 sets)", lambda: ctx.pair_counts(sets))
