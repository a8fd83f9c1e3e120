"""P3 parity (GPU): kmsc_pair_counts through the C ABI vs the oracle's restatement of
GetEdgeWeight / the all-pairs loop (reference lib/core/kmer_set_set.h:158-219).
Bit-exact int64 matrices on the same seeded inputs."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import kmsc
    c = kmsc.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True, params=["hash", "merge", "whash", "lane"])
def build(request, monkeypatch):
    """every case runs with all builds of the main kernel (KMSC_P3_BUILD): the CTA-wide hash table,
    the warp-wide multiway merge, the warp-private hash tables and the lane-private tables give the same
    matrices (whash and lane take up to 64 sets; whash hands a call whose tables fill up back to the other
    builds, lane sends the rows that outgrow a table through its retry kernel)"""
    monkeypatch.setenv("KMSC_P3_BUILD", request.param)
    return request.param


def _mk_sets(ctx, oracle, kmer_sets, K, N, kb):
    import synth
    dev, offs_l, keys_l = [], [], []
    for km in kmer_sets:
        offs, keys = synth.csr_of(km, K, N, kb)
        dev.append(ctx.set_from_csr(K, N, kb, offs, keys))
        offs_l.append(offs)
        keys_l.append(keys)
    return dev, offs_l, keys_l


def _check(ctx, oracle, kmer_sets, K, N, kb, bucket_ids=None):
    dev, offs_l, keys_l = _mk_sets(ctx, oracle, kmer_sets, K, N, kb)
    got, visits = ctx.pair_counts(dev, bucket_ids, with_visits=True)
    import os
    if any(len(k) for k in kmer_sets):
        want_build = {"merge": 1 if len(kmer_sets) <= 128 else 0, "whash": 2 if len(kmer_sets) <= 64 else None, "hash": 0,
                      "lane": 3 if len(kmer_sets) <= 64 and kb <= 4 else None}[os.environ["KMSC_P3_BUILD"]]
        if want_build is not None:
            assert ctx.pair_counts_build() in ((want_build,) if want_build != 2 else (2, 1, 0))   # whash may hand over on overflow
    want, want_visits = oracle.pair_counts(offs_l, keys_l, kb, 1 << N, bucket_ids=bucket_ids, n_threads=8)
    n = len(kmer_sets)
    iu = np.triu_indices(n, 1)
    assert np.array_equal(got[iu], want[iu])
    assert np.array_equal(got, got.T)
    assert visits == want_visits
    sel = np.zeros(1 << N, bool)
    if bucket_ids is None:
        sel[:] = True
    else:
        sel[np.asarray(bucket_ids)] = True
    for i in range(n):
        assert got[i, i] == int(np.diff(offs_l[i])[sel].sum())
    for d in dev:
        d.free()
    return got


CASES = [  # (K, N, key_bytes, n_sets, genome_len)
    (15, 14, 2, 8, 60000),
    (19, 10, 4, 5, 40000),
    (23, 14, 4, 64, 30000),
    (23, 14, 4, 33, 20000),
    (31, 14, 8, 12, 30000),
    (23, 14, 4, 100, 8000),
    (23, 14, 4, 200, 5000),
    (5, 3, 2, 3, 200),
    (23, 14, 4, 300, 3000),   # more than one 256-set accumulator tile: runs over pairs of 128-set groups
]


@pytest.mark.parametrize("K,N,kb,n_sets,glen", CASES)
def test_all_buckets(ctx, oracle, K, N, kb, n_sets, glen):
    import synth
    seqs = synth.phylogeny_sequences(n_sets, glen, p=0.01, seed=K * 1000 + n_sets)
    sets = [synth.kmer_set_of(s, K) for s in seqs]
    _check(ctx, oracle, sets, K, N, kb)


@pytest.mark.parametrize("K,N,kb,n_sets,glen", CASES[:5])
def test_sampled_buckets(ctx, oracle, K, N, kb, n_sets, glen):
    """the reference's own mode: (1<<N)/50 random bucket ids (kmer_set_set.h:123-124)"""
    import synth
    rng = np.random.default_rng(11)
    seqs = synth.phylogeny_sequences(n_sets, glen, p=0.01, seed=7 + n_sets)
    sets = [synth.kmer_set_of(s, K) for s in seqs]
    ids = np.sort(rng.permutation(1 << N)[: (1 << N) // 50]).astype(np.int32)
    _check(ctx, oracle, sets, K, N, kb, ids)
    _check(ctx, oracle, sets, K, N, kb, ids[::-1].copy())  # order must not matter
    _check(ctx, oracle, sets, K, N, kb, np.concatenate([ids, ids[:5]]))  # repeated ids count once


def test_unrelated_sets_and_empty(ctx, oracle):
    """no shared structure (every key distinct) plus empty and tiny sets"""
    rng = np.random.default_rng(5)
    K, N, kb = 23, 14, 4
    sets = [np.unique(rng.integers(0, 1 << 46, 50000, dtype=np.uint64)) for _ in range(6)]
    sets.append(np.zeros(0, np.uint64))
    sets.append(sets[0][:3].copy())
    sets.append(sets[1].copy())
    _check(ctx, oracle, sets, K, N, kb)


def test_skewed_single_bucket(ctx, oracle):
    """everything in one fine bucket: forces the multi-pass (key class) split"""
    rng = np.random.default_rng(6)
    K, N, kb = 23, 14, 4
    base = np.uint64(12345) << np.uint64(32)
    sets = []
    for i in range(10):
        low = np.unique(rng.integers(0, 1 << 20, 20000, dtype=np.uint64))
        sets.append(base | low)
    _check(ctx, oracle, sets, K, N, kb)


def test_all_ones_key(ctx, oracle):
    """key == 0xFFFFFFFF (valid when 2K-N == 32) must not be taken for the empty marker"""
    K, N, kb = 23, 14, 4
    a = np.sort(np.array([0xFFFFFFFF, (5 << 32) | 0xFFFFFFFF, (5 << 32) | 7, (1 << 46) - 1], np.uint64))
    b = np.sort(np.array([0xFFFFFFFF, (5 << 32) | 0xFFFFFFFF, (5 << 32) | 8, (1 << 46) - 1], np.uint64))
    c = np.array([1, (5 << 32) | 0xFFFFFFFE], np.uint64)
    got = _check(ctx, oracle, [a, b, c], K, N, kb)
    assert got[0, 1] == 3 and got[0, 2] == 0


def test_duplicate_keys_multiset(ctx, oracle):
    """GetSampledKmerSet keeps duplicates (kmer_set_compact.h:120-203); the merge then
    counts min multiplicity (kmer_set_set.h:165-180)."""
    import kmsc
    K, N, kb = 23, 14, 4
    a = np.array([5, 5, 5, 9, (3 << 32) | 1, (3 << 32) | 1], np.uint64)
    b = np.array([5, 5, 9, 9, (3 << 32) | 1], np.uint64)
    import synth
    dev, offs_l, keys_l = [], [], []
    for km in (a, b):
        offs, keys = synth.csr_of(km, K, N, kb)
        dev.append(ctx.set_from_csr(K, N, kb, offs, keys))
        offs_l.append(offs)
        keys_l.append(keys)
    got = ctx.pair_counts(dev)
    want, _ = oracle.pair_counts(offs_l, keys_l, kb, 1 << N)
    assert got[0, 1] == want[0, 1] == 4


def _oracle_rows(oracle, offs_l, keys_l, rows, kb, n_buckets, bucket_ids=None):
    """rows of the reference's weight matrix, pair by pair (GetEdgeWeight, kmer_set_set.h:158-184)"""
    n = len(offs_l)
    out = np.zeros((len(rows), n), np.int64)
    for a, r in enumerate(rows):
        for l in range(n):
            w, _ = oracle.pair_counts([offs_l[r], offs_l[l]], [keys_l[r], keys_l[l]], kb, n_buckets, bucket_ids=bucket_ids)
            out[a, l] = w[0, 1]
    return out


@pytest.mark.parametrize("K,N,kb", [(23, 14, 4), (15, 14, 2), (31, 14, 8), (19, 10, 4)])
def test_rows_mode_vs_oracle(ctx, oracle, K, N, kb):
    """row mode = the 3n-2 re-weights after a merge (kmer_set_set.h:385-425), against the oracle's own
    merge counts (not against the product's full matrix): all buckets and a sampled bucket list"""
    import synth
    seqs = synth.phylogeny_sequences(9, 20000, p=0.01, seed=3 + K)
    sets = [synth.kmer_set_of(s, K) for s in seqs]
    sets[4] = sets[4][:50]                      # a tiny set
    sets[6] = np.zeros(0, np.uint64)            # an empty one
    dev, offs_l, keys_l = _mk_sets(ctx, oracle, sets, K, N, kb)
    rows = [2, 5, 8]
    got = ctx.pair_counts_rows(dev, rows)
    want = _oracle_rows(oracle, offs_l, keys_l, rows, kb, 1 << N)
    assert np.array_equal(got, want)
    ids = np.random.default_rng(1).choice(1 << N, max(4, (1 << N) // 50), replace=False).astype(np.int32)
    ids = np.concatenate([ids, ids[:3]])        # an id listed twice counts once
    got = ctx.pair_counts_rows(dev, rows, bucket_ids=ids)
    want = _oracle_rows(oracle, offs_l, keys_l, rows, kb, 1 << N, bucket_ids=np.unique(ids))
    assert np.array_equal(got, want)
    # more rows than one launch takes, a row listed twice, a row against itself
    rows = [0, 1, 2, 3, 4, 5, 6, 7, 8, 3]
    got = ctx.pair_counts_rows(dev, rows)
    assert np.array_equal(got, _oracle_rows(oracle, offs_l, keys_l, rows, kb, 1 << N))


def test_rows_mode_dense_and_duplicates(ctx, oracle):
    """runs far longer than the staging tile (one bucket holds everything) and multiset inputs:
    the merge counts min multiplicity like the reference's loop"""
    K, N, kb = 23, 14, 4
    rng = np.random.default_rng(5)
    base = (np.uint64(77) << np.uint64(32)) | np.unique(rng.integers(0, 1 << 20, 30000).astype(np.uint64))
    a = base[rng.random(len(base)) < 0.7]
    b = base[rng.random(len(base)) < 0.7]
    c = np.sort(np.concatenate([a[:2000], a[:2000], b[:10]]))   # duplicates
    dev, offs_l, keys_l = [], [], []
    import synth
    for km in (a, b, c):
        offs, keys = synth.csr_of(km, K, N, kb)
        dev.append(ctx.set_from_csr(K, N, kb, offs, keys))
        offs_l.append(offs)
        keys_l.append(keys)
    got = ctx.pair_counts_rows(dev, [0, 2])
    assert np.array_equal(got, _oracle_rows(oracle, offs_l, keys_l, [0, 2], kb, 1 << N))


def test_set_roundtrip_size_hash(ctx, oracle):
    import synth
    for K, N, kb in [(15, 14, 2), (23, 14, 4), (31, 14, 8), (19, 10, 4)]:
        km = synth.kmer_set_of(synth.random_genome(30000, K), K)
        offs, keys = synth.csr_of(km, K, N, kb)
        s = ctx.set_from_csr(K, N, kb, offs, keys)
        o2, k2 = s.to_csr()
        assert np.array_equal(o2, offs) and np.array_equal(k2, keys)
        assert s.Size() == len(km) and s.Hash() == oracle.set_hash(km)
        s2 = ctx.set_from_kmers(K, N, kb, km)
        assert np.array_equal(s2.to_kmers(), km) and s2.Hash() == s.Hash()
