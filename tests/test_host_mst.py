"""Host logic without a device: the spanning tree of the `mst` driver in the C++ facade (MstTree, host/kmsc/kmer_set_set.h)
against the oracle's restatement kmsc_o_mst (union-find of reference lib/core/parallel_disjoint_set.h:53-106)
on random intersection matrices with many tied distances: identical edge lists, in order, with distances."""
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "kmer-sets-compression_b200" / "host"


def _matrices():
    rng = np.random.default_rng(12)
    out = []
    for n in (1, 2, 3, 5, 8, 17, 40, 64):
        for ties in (False, True, True):
            sizes = rng.integers(50, 60 if ties else 5000, n)
            W = np.zeros((n, n), np.int64)
            for i in range(n):
                W[i, i] = sizes[i]
                for j in range(i + 1, n):
                    hi = int(min(sizes[i], sizes[j]))
                    W[i, j] = W[j, i] = int(rng.integers(hi - 3 if ties else 0, hi + 1))
            out.append(W)
    # identical twins and a star of equal distances
    W = np.full((6, 6), 10, np.int64)
    out.append(W)
    W = np.full((7, 7), 0, np.int64)
    np.fill_diagonal(W, 100)
    out.append(W)
    return out


def test_mst_tree_vs_oracle(oracle):
    r = subprocess.run(["make", "-s", "-C", str(HOST), "bin/mst_test"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    mats = _matrices()
    text = "".join(f"{len(W)}\n" + "\n".join(" ".join(str(int(x)) for x in row) for row in W) + "\n" for W in mats)
    r = subprocess.run([str(HOST / "bin" / "mst_test")], input=text, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    at = 0
    for W in mats:
        n = len(W)
        head = lines[at].split()
        assert head[0] == "tree" and int(head[1]) == n
        ne = int(head[2])
        got = [tuple(int(x) for x in lines[at + 1 + k].split()) for k in range(ne)]
        at += 1 + ne
        edges, dist = oracle.mst(W)
        want = [(int(a), int(b), int(d)) for (a, b), d in zip(edges, dist)]
        assert got == want, (n, got[:5], want[:5])
        assert ne == max(0, n - 1)
