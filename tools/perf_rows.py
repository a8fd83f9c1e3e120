"""Row mode against the full matrix: kmsc_pair_counts_rows (3 rows, the re-weights of one greedy iteration,
reference lib/core/kmer_set_set.h:385-425) vs kmsc_pair_counts on the same device sets. GPU only.
  SETS=64 KMERS=10000000 python tools/perf_rows.py"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import kmsc
import bench

K, N, KB = 23, 14, 4
n_sets = int(os.environ.get("SETS", "64"))
kmers = int(os.environ.get("KMERS", "10000000"))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = kmsc.Context(0, stream.cuda_stream)
G = kmers + K - 1
seqs = bench.gen_sequences_torch(n_sets, G, 0.002, dev)
str_offs = np.array([0, G], np.int64)
pinned = []
for s in seqs:
    w = bench.pack_torch(s)
    h = torch.empty(w.numel(), dtype=torch.int64, pin_memory=True)
    h.copy_(w)
    pinned.append(h)
del seqs
sets = ctx.sets_from_packed_batch(K, N, KB, None, [str_offs] * n_sets, words_ptrs=[h.data_ptr() for h in pinned])
rows = [1, n_sets // 2, n_sets - 1]
ids = np.sort(np.random.default_rng(1).choice(1 << N, (1 << N) // 50, replace=False)).astype(np.int32)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r


t_full, W = timeit(lambda: ctx.pair_counts(sets))
t_rows, R = timeit(lambda: ctx.pair_counts_rows(sets, rows))
assert np.array_equal(R, W[rows]), "row mode differs from the matrix rows"
t_full_s, Ws = timeit(lambda: ctx.pair_counts(sets, bucket_ids=ids))
t_rows_s, Rs = timeit(lambda: ctx.pair_counts_rows(sets, rows, bucket_ids=ids))
assert np.array_equal(Rs, Ws[rows])
keys = sum(s.n_keys for s in sets)
print(f"{n_sets} sets x {kmers}: full matrix {t_full:.3f} ms, 3 rows {t_rows:.3f} ms ({keys * KB / t_rows / 1e6:.0f} GB/s of column keys, "
      f"{t_full / t_rows:.2f} x cheaper); 2% bucket sample: full {t_full_s:.3f} ms, 3 rows {t_rows_s:.3f} ms", flush=True)
