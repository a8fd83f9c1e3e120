"""Seeded synthetic inputs of the shapes BASELINE.json names:
a random genome and k-mer sets derived from it by random substitution. numpy only;
this is bench/test scaffolding, not part of the hot path."""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", np.uint8)


def random_genome(n: int, seed: int = 12345) -> np.ndarray:
    """i.i.d. uniform bases as codes 0..3 (A C G T)."""
    return np.random.default_rng(seed).integers(0, 4, n, dtype=np.uint8)


def mutate(codes: np.ndarray, p: float, seed: int) -> np.ndarray:
    """per-base substitution with probability p (always to a different base)."""
    rng = np.random.default_rng(seed)
    out = codes.copy()
    m = rng.random(len(codes)) < p
    out[m] = (out[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
    return out


def to_ascii(codes: np.ndarray) -> bytes:
    return BASES[codes].tobytes()


def kmers_of(codes: np.ndarray, K: int, canonical: bool = True) -> np.ndarray:
    """2K-bit k-mer value at every position (reference lib/core/kmer.h:22-46, 103-133)."""
    n = len(codes) - K + 1
    if n <= 0:
        return np.zeros(0, np.uint64)
    c = codes.astype(np.uint64)
    fwd = np.zeros(n, np.uint64)
    rc = np.zeros(n, np.uint64)
    for i in range(K):
        fwd = (fwd << np.uint64(2)) | c[i:i + n]
        rc |= (np.uint64(3) - c[i:i + n]) << np.uint64(2 * i)
    return np.minimum(fwd, rc) if canonical else fwd


def kmer_set_of(codes: np.ndarray, K: int, canonical: bool = True) -> np.ndarray:
    """ascending distinct k-mer values = the KmerSet the sequence spells."""
    return np.unique(kmers_of(codes, K, canonical))


def csr_of(kmers: np.ndarray, K: int, N: int, key_bytes: int):
    """ascending k-mers -> (offs int64[2^N+1], keys) per lib/core/kmer_set.h:22-31."""
    kb = 2 * K - N
    buckets = (kmers >> np.uint64(kb)).astype(np.int64)
    offs = np.zeros((1 << N) + 1, np.int64)
    np.cumsum(np.bincount(buckets, minlength=1 << N), out=offs[1:])
    dt = {2: np.uint16, 4: np.uint32, 8: np.uint64}[key_bytes]
    keys = (kmers & np.uint64((1 << kb) - 1)).astype(dt)
    return offs, keys


def phylogeny_sequences(n_sets: int, genome_len: int, p: float = 0.002, seed: int = 12345):
    """C2/C3 shape: set 0 = genome, set i>=1 = set (i-1)//2 with substitution prob p
    (seed 1000+i): a binary phylogeny, so overlaps are hierarchical."""
    seqs = [random_genome(genome_len, seed)]
    for i in range(1, n_sets):
        seqs.append(mutate(seqs[(i - 1) // 2], p, 1000 + i))
    return seqs


def window_sequences(n_sets: int, genome_len: int, window: int, p: float = 0.005, seed: int = 12345):
    """C1 shape: set i = window at offset floor(i*(G-window)/(n-1)) with substitution prob p."""
    g = random_genome(genome_len, seed)
    out = []
    for i in range(n_sets):
        off = (i * (genome_len - window)) // max(1, n_sets - 1)
        out.append(mutate(g[off:off + window], p, 1000 + i))
    return out


def split_strings(codes: np.ndarray, K: int, piece: int = 10000):
    """cut a sequence into overlapping pieces (overlap K-1) spelling the same k-mers:
    an SPSS-like multi-string form of one sequence."""
    out = []
    i = 0
    n = len(codes)
    while i + K <= n:
        j = min(n, i + piece)
        out.append(to_ascii(codes[i:j]))
        if j == n:
            break
        i = j - (K - 1)
    return out
