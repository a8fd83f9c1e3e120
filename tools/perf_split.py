"""P4 timing probe: kmsc_pair_split_batch on the edges of a small phylogeny at C2 density
(10 M canonical 23-mers per set), for several tile fill factors. GPU only."""
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
n_sets = int(os.environ.get("SETS", "16"))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = kmsc.Context(0, stream.cuda_stream)
G = 10_000_000 + K - 1
seqs = bench.gen_sequences_torch(n_sets, G, 0.002, dev)
str_offs = np.array([0, G], np.int64)
sets = []
for s in seqs:
    w = bench.pack_torch(s).cpu()
    sets.append(ctx.set_from_packed(K, N, KB, None, str_offs, words_ptr=w.data_ptr()))
W = ctx.pair_counts(sets)
edges = [((i - 1) // 2, i) for i in range(1, n_sets)]
js, ks = [sets[a] for a, _ in edges], [sets[b] for _, b in edges]
hint = np.array([W[a, b] for a, b in edges], np.int64)
in_bytes = sum(a.n_keys + b.n_keys for a, b in zip(js, ks)) * KB


def run(label, reps=5, **kw):
    for _ in range(2):
        r = ctx.pair_split_batch(js, ks, **kw)
        for l in r:
            for s in (l or []):
                s.free()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        r = ctx.pair_split_batch(js, ks, **kw)
        for l in r:
            for s in (l or []):
                s.free()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    wall = (time.perf_counter() - t0) / reps * 1e3
    print(f"{label:50s} {ms:8.3f} ms/batch  wall {wall:8.3f}  {ms / len(edges) * 1e3:7.1f} us/pair  in {in_bytes / ms / 1e6:7.1f} GB/s", flush=True)


for fill in os.environ.get("FILLS", "0.3,0.4,0.45,0.5,0.55").split(","):
    os.environ["KMSC_SPLIT_FILL"] = fill
    run(f"fill={fill} hint, diffs only", inter_hint=hint, want_inter=False)
os.environ.pop("KMSC_SPLIT_FILL")
if os.environ.get("ONLY"):
    sys.exit(0)
run("default hint, all three outputs", inter_hint=hint)
run("default no hint, diffs only", want_inter=False)
run("default no hint, all three", )
run("default hint, no outputs (count)", inter_hint=hint, want_inter=False, want_j=False, want_k=False)
