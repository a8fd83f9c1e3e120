"""Quick P3 timing probe (not the bench): n sets x m k-mers, wall clock around the
synchronous C-ABI call, checked against the oracle on a bucket sample."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200")); sys.path.insert(0, str(ROOT / "tests"))
import kmsc, synth
from _oracle import Oracle

n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 64
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
K, N, kb = 23, 14, 4
t0 = time.time()
seqs = synth.phylogeny_sequences(n_sets, glen, p=0.002)
ctx = kmsc.Context(0)
dev, offs_l, keys_l = [], [], []
for s in seqs:
    km = synth.kmer_set_of(s, K)
    offs, keys = synth.csr_of(km, K, N, kb)
    dev.append(ctx.set_from_csr(K, N, kb, offs, keys)); offs_l.append(offs); keys_l.append(keys)
tot = sum(d.n_keys for d in dev)
print(f"gen {time.time()-t0:.1f}s  total keys {tot}  bytes {tot*kb/1e9:.3f} GB", flush=True)
for it in range(6):
    t = time.time(); W, visits = ctx.pair_counts(dev, with_visits=True); dt = time.time() - t
    print(f"iter {it}: {dt*1e3:.3f} ms  key-visits/s {visits/dt:.3e}  GB/s {tot*kb/dt/1e9:.1f}", flush=True)
ids = np.arange(0, 1 << N, 64, dtype=np.int32)
t = time.time(); Ws = ctx.pair_counts(dev, ids); print("sampled call", (time.time()-t)*1e3, "ms")
o = Oracle()
want, _ = o.pair_counts(offs_l, keys_l, kb, 1 << N, bucket_ids=ids, n_threads=16)
iu = np.triu_indices(n_sets, 1)
print("sampled parity:", np.array_equal(Ws[iu], want[iu]))
t = time.time(); want_full, v = o.pair_counts(offs_l[:8], keys_l[:8], kb, 1 << N, n_threads=8); dt = time.time() - t
print("full parity (first 8 sets):", np.array_equal(W[:8, :8][np.triu_indices(8, 1)], want_full[np.triu_indices(8, 1)]),
      f"cpu {dt:.2f}s {v/dt:.3e} key-visits/s (8 threads)")
