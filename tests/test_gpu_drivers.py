"""Driver-level parity (GPU): what `kmerset-multiple-compress` decides, not only what its kernels compute.

 * greedy driver (reference lib/core/kmer_set_set.h:109-427) replayed on the fixtures the seeded,
   single-worker reference produced (tests/golden/ref_golden.json: SPSS, its bucket sample, its merge
   list): the initial weight matrix must equal the oracle's GetEdgeWeight matrix over the same
   buckets, the first merge must be the reference's (:308-316), every later merge must agree for as
   long as both ran (the stop rule reads SPSS text weights, which differ), every input set must
   reconstruct, and the reference's OWN KmerSetSetReader (:672-755, oracle/_ref) must read the
   directory this repository dumped.
 * mst driver (north_star): tree edges in order, distances and the (size, XOR hash)
   of both difference sets of every edge equal oracle.mst / oracle set algebra on inputs with tied
   distances; the dumped tree reconstructs every set.
"""
import json
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "kmer-sets-compression_b200" / "host"
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
from _oracle import CONFIGS, Ref  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).parent / "golden" / "ref_golden.json").read_text())


@pytest.fixture(scope="module")
def bins():
    r = subprocess.run(["make", "-s", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return HOST / "bin"


def parse_trace(path):
    t = {"merges": [], "edges": []}
    for line in Path(path).read_text().split("\n"):
        p = line.split()
        if not p:
            continue
        if p[0] == "driver":
            t["driver"] = p[1]
        elif p[0] == "n":
            t["n"] = int(p[1])
        elif p[0] == "W":
            t["W"] = np.array([int(x) for x in p[1:]], np.int64).reshape(t["n"], t["n"])
        elif p[0] == "merge":
            t["merges"].append([int(x) for x in p[1:]])
        elif p[0] == "edge":
            t["edges"].append([int(x) for x in p[1:]])
    return t


def decompress(bins, K, outdir, n):
    r = subprocess.run([str(bins / "kmerset-multiple-decompress"), f"--k={K}", f"--n={n}", str(outdir)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    hashes = [int(x) for x in re.findall(r"kmer_set.Hash\(\) = (\d+)", r.stderr)]
    sizes = [int(x) for x in re.findall(r"kmer_set.Size\(\) = (\d+)", r.stderr)]
    return list(zip(sizes, hashes))


@pytest.mark.parametrize("idx", range(len(GOLD["kmersetset"])))
def test_greedy_driver_vs_seeded_reference(bins, oracle, tmp_path, idx):
    e = GOLD["kmersetset"][idx]
    K, N, kb = CONFIGS[e["cfg"]]
    files = []
    for i, spss in enumerate(e["spss"]):
        f = tmp_path / f"s{i}.txt"
        f.write_text("\n".join(spss) + "\n")
        files.append(str(f))
    ids_file = tmp_path / "ids.txt"
    ids_file.write_text("\n".join(str(x) for x in e["bucket_ids"]) + "\n")
    outdir, trace = tmp_path / "dump", tmp_path / "trace.txt"
    r = subprocess.run([str(bins / "kmerset-multiple-compress"), f"--k={K}", f"--out={outdir}", f"--bucket_ids_file={ids_file}",
                        f"--trace={trace}", f"--max_iterations={len(e['merges'])}"] + files,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    t = parse_trace(trace)
    n = len(files)
    # the weight matrix over the reference's own bucket sample
    ids = np.array(e["bucket_ids"], np.int32)
    offs_l, keys_l = [], []
    for spss in e["spss"]:
        offs, keys = oracle.sampled_set(spss, K, N, True, ids)
        offs_l.append(offs)
        keys_l.append(keys)
    want_w, _ = oracle.pair_counts(offs_l, keys_l, 8, len(ids), bucket_ids=np.arange(len(ids), dtype=np.int32))
    iu = np.triu_indices(n, 1)
    assert np.array_equal(t["W"][iu], want_w[iu])
    # the merge sequence: identical to the seeded reference's for as long as both ran
    assert t["merges"], "the driver made no merge"
    assert t["merges"][0] == e["merges"][0]
    common = min(len(t["merges"]), len(e["merges"]))
    assert t["merges"][:common] == e["merges"][:common]
    assert common >= min(len(e["merges"]), n // 8 + 1), "stopped before the first interval check"
    # every input reconstructs (reference test/kmer_set_set.cc:15-35), through our reader ...
    want = list(zip(e["in_sizes"], e["in_hashes"]))
    assert decompress(bins, K, outdir, n) == want
    # ... and through the reference's own KmerSetSetReader pointed at OUR directory
    if Ref.available():
        ref = Ref()
        for i in range(n):
            size, h, n_nodes = ref.reader_get(e["cfg"], outdir, True, i)
            assert (size, h) == want[i]
        assert n_nodes == int((outdir / "meta.txt").read_text().split("\n")[1])


def _mst_inputs(seed, K):
    """small related sets with tied symmetric-difference distances: identical twins (d = 0 twice) and
    two children mutated at the same number of non-overlapping places"""
    import synth
    rng = np.random.default_rng(seed)
    base = synth.random_genome(3000 + K - 1, seed)
    seqs = [base, base.copy()]                       # 0 == 1
    for lo in (100, 900, 1700):                      # three children, disjoint edits of equal size -> ties
        s = base.copy()
        for p in range(lo, lo + 200, 40):
            s[p] = (s[p] + 1) & 3
        seqs.append(s)
    seqs.append(seqs[2].copy())                      # 5 == 2
    g = seqs[3].copy()
    g[2500] = (g[2500] + 2) & 3
    seqs.append(g)
    seqs.append(synth.mutate(base, 0.02, seed + 7))
    order = rng.permutation(len(seqs))
    return [seqs[i] for i in order]


@pytest.mark.parametrize("K,seed", [(15, 1), (23, 2), (19, 3), (23, 4)])
def test_mst_driver_vs_oracle(bins, oracle, tmp_path, K, seed):
    import synth
    N, kb = {15: (14, 2), 19: (10, 4), 23: (14, 4)}[K]
    seqs = _mst_inputs(seed, K)
    n = len(seqs)
    files, sets = [], []
    for i, s in enumerate(seqs):
        f = tmp_path / f"s{i}.txt"
        f.write_bytes(b"\n".join(synth.split_strings(s, K, 700)) + b"\n")
        files.append(str(f))
        sets.append(synth.kmer_set_of(s, K, True))
    outdir, trace = tmp_path / "dump", tmp_path / "trace.txt"
    r = subprocess.run([str(bins / "kmerset-multiple-compress"), f"--k={K}", "--driver=mst", f"--out={outdir}",
                        f"--trace={trace}"] + files, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    t = parse_trace(trace)
    assert t["driver"] == "mst"
    # exact matrix
    csr = [oracle.to_csr(s, K, N, kb) for s in sets]
    want_w, _ = oracle.pair_counts([c[0] for c in csr], [c[1] for c in csr], kb, 1 << N)
    want_w = want_w + want_w.T + np.diag([len(s) for s in sets])
    assert np.array_equal(t["W"], want_w)
    # the tree: same edges in the same order, same distances; ties exist in this input
    edges, dist = oracle.mst(want_w)
    d_all = want_w.diagonal()[:, None] + want_w.diagonal()[None, :] - 2 * want_w
    assert len(np.unique(d_all[np.triu_indices(n, 1)])) < n * (n - 1) // 2, "input has no tied distances"
    assert [[p, c, d] for (p, c), d in zip(edges.tolist(), dist.tolist())] == [x[:3] for x in t["edges"]]
    # both difference sets of every edge
    for (p, c), x in zip(edges.tolist(), t["edges"]):
        add = oracle.set_sub(sets[c], sets[p])
        dele = oracle.set_sub(sets[p], sets[c])
        assert x[3:] == [len(add), oracle.set_hash(add), len(dele), oracle.set_hash(dele)]
    # the dumped tree reconstructs every set
    want = [(len(s), oracle.set_hash(s)) for s in sets]
    assert decompress(bins, K, outdir, n) == want
