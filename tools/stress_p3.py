"""Repeat the P3 edge cases that force table overflows / key-class splits, to flush out
timing-dependent bugs. Prints one line per iteration; exits non-zero on a mismatch."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200")); sys.path.insert(0, str(ROOT / "tests"))
import kmsc, synth
from _oracle import Oracle

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
o = Oracle()
ctx = kmsc.Context(0)
K, N, kb = 23, 14, 4
rng = np.random.default_rng(6)
base = np.uint64(12345) << np.uint64(32)
sets = [base | np.unique(rng.integers(0, 1 << 20, 20000, dtype=np.uint64)) for _ in range(10)]
dev, offs_l, keys_l = [], [], []
for km in sets:
    offs, keys = synth.csr_of(km, K, N, kb)
    dev.append(ctx.set_from_csr(K, N, kb, offs, keys)); offs_l.append(offs); keys_l.append(keys)
want, _ = o.pair_counts(offs_l, keys_l, kb, 1 << N, n_threads=8)
iu = np.triu_indices(len(sets), 1)
bad = 0
for it in range(iters):
    t = time.time()
    try:
        got = ctx.pair_counts(dev)
        ok = np.array_equal(got[iu], want[iu])
        st = ctx.pair_counts_stats()
        print(f"iter {it}: {'ok' if ok else 'MISMATCH'} {1e3*(time.time()-t):.1f} ms retries={st.get('retries')} distinct={st.get('distinct')}", flush=True)
        bad += 0 if ok else 1
    except Exception as e:
        print(f"iter {it}: ERROR {e} after {1e3*(time.time()-t):.1f} ms", flush=True)
        bad += 1
sys.exit(1 if bad else 0)
