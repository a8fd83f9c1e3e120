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

# keys equal to the table's empty marker (valid when 2K - N == 32), alone and next to ordinary
# keys, in one and in several buckets, alternating between shapes inside one context
E = 0xFFFFFFFF
cases = [
    [[E], [E]],
    [[(5 << 32) | E, (5 << 32) | 7], [(5 << 32) | E, (5 << 32) | 8]],
    [[E, (5 << 32) | E, (5 << 32) | 7, (1 << 46) - 1], [E, (5 << 32) | E, (5 << 32) | 8, (1 << 46) - 1], [1, (5 << 32) | (E - 1)]],
    [[E, 7, 8, 9], [E, 7], [E], [7]],
]
prepared = []
for sets_ in cases:
    d_, o_, k_ = [], [], []
    for km in sets_:
        offs, keys = synth.csr_of(np.array(sorted(km), np.uint64), K, N, kb)
        d_.append(ctx.set_from_csr(K, N, kb, offs, keys)); o_.append(offs); k_.append(keys)
    w_, _ = o.pair_counts(o_, k_, kb, 1 << N, n_threads=2)
    prepared.append((d_, w_))
for it in range(iters * 5):
    d_, w_ = prepared[it % len(prepared)]
    got = ctx.pair_counts(d_)
    iu2 = np.triu_indices(len(d_), 1)
    if not (np.array_equal(got[iu2], w_[iu2]) and all(got[i, i] == d_[i].n_keys for i in range(len(d_)))):
        print(f"special-key case {it % len(prepared)} iter {it}: MISMATCH {got.tolist()}", flush=True)
        bad += 1
print("special-key cases:", "ok" if not bad else f"{bad} failures")
sys.exit(1 if bad else 0)
