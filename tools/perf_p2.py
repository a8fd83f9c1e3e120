"""P2 timing probe: kmsc_sets_from_packed_batch on C2-shaped sets (10 M canonical 23-mers each)
from pinned host memory; wall clock and CUDA-event time per call. GPU only.
  SETS=64 REPS=3 python tools/perf_p2.py        (ONCE=1: a single call, for a profiler run)"""
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
reps = int(os.environ.get("REPS", "3"))
kmers = int(os.environ.get("KMERS", "10000000"))
once = os.environ.get("ONCE") == "1"
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
torch.cuda.synchronize()


def call():
    return ctx.sets_from_packed_batch(K, N, KB, None, [str_offs] * n_sets, words_ptrs=[h.data_ptr() for h in pinned])


if os.environ.get("CHECK") == "1":
    # the staged sort against the general pipeline on the same input: sizes and XOR hashes
    ss = call()
    os.environ["KMSC_P2_LEGACY"] = "1"
    for i, h in enumerate(pinned[:4]):
        ref = ctx.set_from_packed(K, N, KB, None, str_offs, words_ptr=h.data_ptr())
        assert (ref.n_keys, ref.Hash()) == (ss[i].n_keys, ss[i].Hash()), f"set {i}: staged sort and general pipeline differ"
        a, b = ref.to_csr(), ss[i].to_csr()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        ref.free()
    del os.environ["KMSC_P2_LEGACY"]
    print(f"CHECK ok: {n_sets} x {kmers}, first sets identical to the general pipeline", flush=True)
    for s in ss:
        s.free()

if once:
    ss = call()
    torch.cuda.synchronize()
    print("n_keys", ss[0].n_keys)
    sys.exit(0)
for _ in range(2):
    for s in call():
        s.free()
torch.cuda.synchronize()
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    ss = call()
    e1.record(stream)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"batch of {n_sets} x {kmers}: wall {1e3 * (t1 - t0):.2f} ms, events {e0.elapsed_time(e1):.2f} ms, "
          f"{1e3 * (t1 - t0) / n_sets:.3f} ms per set; launches so far {ctx.launch_count()}", flush=True)
    for s in ss:
        s.free()
# the same sets one by one through kmsc_set_from_packed
t0 = time.perf_counter()
ss = [ctx.set_from_packed(K, N, KB, None, str_offs, words_ptr=h.data_ptr()) for h in pinned]
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"one by one: wall {1e3 * (t1 - t0):.2f} ms, {1e3 * (t1 - t0) / n_sets:.3f} ms per set")
