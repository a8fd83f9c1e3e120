"""BASELINE configs[1] THROUGH THE EXECUTABLE: 64 sets x 10 M canonical 23-mers as SPSS text files ->
kmerset-multiple-compress --driver=mst (load, batched decode, exact matrix, spanning tree, both difference
sets of every edge, SPSS of all of them, dump) -> kmerset-multiple-decompress of a few sets. Checks, outside
the timed runs: the tree against the oracle's Kruskal restatement on the traced matrix, matrix entries and the
difference sets of some edges against numpy k-mer sets of the same sequences, the decompressed sets against the
originals (size, XOR hash). usage: c2_exec.py [n_sets] [kmers]"""
import os, re, shutil, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200")); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench, synth
from _oracle import Oracle

n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kmers = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
K = 23
dev = torch.device("cuda", 0)
seqs = bench.gen_sequences_torch(n_sets, kmers + K - 1, 0.002, dev)
codes = [s.cpu().numpy() for s in seqs]
del seqs
torch.cuda.empty_cache()
tmp = tempfile.mkdtemp(prefix="kmsc_c2x_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
t = time.time()
files = []
for i, c in enumerate(codes):
    f = os.path.join(tmp, f"s{i}.txt")
    bench.write_spss(c, f)
    files.append(f)
print(f"{n_sets} SPSS files, {sum(os.path.getsize(f) for f in files) / 1e6:.0f} MB, written in {time.time() - t:.1f} s", flush=True)
bins = ROOT / "kmer-sets-compression_b200" / "host" / "bin"
out, trace = os.path.join(tmp, "dump"), os.path.join(tmp, "trace.txt")
t = time.time()
n_gpus = int(os.environ.get("GPUS", "1"))
r = subprocess.run([str(bins / "kmerset-multiple-compress"), f"--k={K}", "--driver=mst", f"--workers={os.cpu_count()}", f"--out={out}",
                    f"--trace={trace}"] + ([f"--gpus={n_gpus}"] if n_gpus > 1 else []) + files, capture_output=True, text=True)
wall = time.time() - t
print(f"kmerset-multiple-compress --driver=mst{f' --gpus={n_gpus}' if n_gpus > 1 else ''}: rc={r.returncode}, wall {wall:.2f} s")
for l in r.stderr.split("\n"):
    if "phases" in l or "seconds" in l or "edges" in l or "rror" in l:
        print("   ", l)
assert r.returncode == 0
# ---- checks ----
o = Oracle()
lines = open(trace).read().split("\n")
W = np.array(lines[2].split()[1:], dtype=np.int64).reshape(n_sets, n_sets)
edges = [tuple(int(x) for x in l.split()[1:]) for l in lines if l.startswith("edge ")]
want_edges, want_dist = o.mst(W)
assert [(e[0], e[1], e[2]) for e in edges] == [(int(a), int(b), int(d)) for (a, b), d in zip(want_edges, want_dist)], "tree differs from the oracle's"
ksets = {}
def kset(i):
    if i not in ksets:
        ksets[i] = synth.kmer_set_of(codes[i], K, True)
    return ksets[i]
chk = 0
for e in edges[:2] + edges[-1:]:
    p, c = e[0], e[1]
    a, b = kset(p), kset(c)
    inter = np.intersect1d(a, b, assume_unique=True)
    assert W[p, c] == len(inter), (p, c)
    add, dele = np.setdiff1d(b, a, assume_unique=True), np.setdiff1d(a, b, assume_unique=True)
    assert (e[3], e[4]) == (len(add), o.set_hash(add)) and (e[5], e[6]) == (len(dele), o.set_hash(dele)), (p, c)
    chk += 1
print(f"tree = oracle Kruskal on the traced matrix ({len(edges)} edges); {chk} edges: |S_p & S_c| and both difference sets (size, hash) = numpy")
t = time.time()
r = subprocess.run([str(bins / "kmerset-multiple-decompress"), f"--k={K}", f"--n={n_sets}", out], capture_output=True, text=True)
print(f"kmerset-multiple-decompress of all {n_sets} sets: rc={r.returncode}, wall {time.time() - t:.2f} s")
assert r.returncode == 0, r.stderr[-500:]
hashes = [int(x) for x in re.findall(r"kmer_set.Hash\(\) = (\d+)", r.stderr)]
sizes = [int(x) for x in re.findall(r"kmer_set.Size\(\) = (\d+)", r.stderr)]
for i in sorted(ksets):
    assert (sizes[i], hashes[i]) == (len(ksets[i]), o.set_hash(ksets[i])), i
print(f"decompressed sets {sorted(ksets)}: size and XOR hash equal the originals; dump = {sum(f.stat().st_size for f in Path(out).iterdir()) / 1e6:.0f} MB "
      f"for {sum(os.path.getsize(f) for f in files) / 1e6:.0f} MB of input")
# ---- the reference's own driver (greedy merges over sampled weights, lib/core/kmer_set_set.h:109-427) on the same files
if len(sys.argv) > 3 and sys.argv[3] == "greedy":
    out2 = os.path.join(tmp, "dump_greedy")
    t = time.time()
    r = subprocess.run([str(bins / "kmerset-multiple-compress"), f"--k={K}", "--seed=11", f"--workers={os.cpu_count()}", f"--out={out2}"] + files,
                       capture_output=True, text=True, env=dict(os.environ, KMSC_TIMING="1"))
    print(f"kmerset-multiple-compress (greedy, 2 % bucket sample): rc={r.returncode}, wall {time.time() - t:.2f} s")
    for l in r.stderr.split("\n"):
        if "timing" in l or "seconds" in l or "rror" in l:
            print("   ", l)
    assert r.returncode == 0
    t = time.time()
    r = subprocess.run([str(bins / "kmerset-multiple-decompress"), f"--k={K}", f"--n={n_sets}", out2], capture_output=True, text=True)
    print(f"kmerset-multiple-decompress of all {n_sets} sets: rc={r.returncode}, wall {time.time() - t:.2f} s")
    assert r.returncode == 0, r.stderr[-500:]
    hashes = [int(x) for x in re.findall(r"kmer_set.Hash\(\) = (\d+)", r.stderr)]
    sizes = [int(x) for x in re.findall(r"kmer_set.Size\(\) = (\d+)", r.stderr)]
    for i in sorted(ksets):
        assert (sizes[i], hashes[i]) == (len(ksets[i]), o.set_hash(ksets[i])), i
    print(f"greedy dump: decompressed sets {sorted(ksets)} equal the originals; {sum(f.stat().st_size for f in Path(out2).iterdir()) / 1e6:.0f} MB")
shutil.rmtree(tmp)
