"""Repeat P3 cases with the merge build against the hash build's matrix (same library call,
KMSC_P3_BUILD switched between calls), to flush out timing-dependent bugs."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200")); sys.path.insert(0, str(ROOT / "tests"))
import kmsc, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = kmsc.Context(0)
CASES = [(15, 14, 2, 8, 60000, 0.01), (23, 14, 4, 64, 30000, 0.01), (31, 14, 8, 12, 30000, 0.01), (23, 14, 4, 100, 8000, 0.01),
         (19, 10, 4, 5, 40000, 0.01), (23, 14, 4, 16, 2000000, 0.002)]
bad = 0
for (K, N, kb, n, glen, p) in CASES:
    seqs = synth.phylogeny_sequences(n, glen, p=p, seed=K * 1000 + n)
    dev = []
    for s in seqs:
        offs, keys = synth.csr_of(synth.kmer_set_of(s, K), K, N, kb)
        dev.append(ctx.set_from_csr(K, N, kb, offs, keys))
    os.environ["KMSC_P3_BUILD"] = "hash"
    want = ctx.pair_counts(dev)
    os.environ["KMSC_P3_BUILD"] = "merge"
    for it in range(iters):
        try:
            got = ctx.pair_counts(dev)
            ok = np.array_equal(got, want) and ctx.pair_counts_build() == 1
            if not ok:
                d = np.argwhere(got != want)
                print(f"case {(K, N, kb, n, glen)} iter {it}: MISMATCH at {d[:4].tolist()} got {got[tuple(d[0])]} want {want[tuple(d[0])]}", flush=True)
                bad += 1
        except Exception as e:
            print(f"case {(K, N, kb, n, glen)} iter {it}: ERROR {e}", flush=True)
            bad += 1
            sys.exit(2)
    st = ctx.pair_counts_stats()
    print(f"case {(K, N, kb, n, glen)}: {iters} iterations, bad so far {bad}, stats {st}", flush=True)
    for d in dev:
        d.free()
print("merge stress:", "ok" if not bad else f"{bad} failures")
sys.exit(1 if bad else 0)
