"""Live comparison of the C restatement against the reference's own code
(oracle/_ref/libkmsc_ref.so). Skipped where _ref was not prebuilt."""
import numpy as np
import pytest

from _oracle import CONFIGS


def _randseq(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5])
def test_kmer_ops(oracle, ref, cfg):
    K, N, _ = CONFIGS[cfg]
    rng = np.random.default_rng(cfg)
    for _ in range(50):
        s = _randseq(rng, K)
        b = oracle.bits(s)
        assert ref.kmer_op(cfg, 0, s=s)[0] == b
        assert ref.kmer_op(cfg, 1, bits=b)[0] == oracle.complement(b, K)
        assert ref.kmer_op(cfg, 2, bits=b)[0] == oracle.canonical(b, K)
        c = str(rng.choice(list("ACGT")))
        assert ref.kmer_op(cfg, 3, bits=b, c=c)[0] == oracle.next(b, K, c)
        assert ref.kmer_op(cfg, 4, bits=b, c=c)[0] == oracle.prev(b, K, c)
        assert ref.bucket_key(cfg, b)[:2] == oracle.bucket_key(b, K, N)


@pytest.mark.parametrize("cfg,n_workers", [(2, 1), (4, 4), (5, 2)])
def test_counter(oracle, ref, cfg, n_workers):
    K = CONFIGS[cfg][0]
    rng = np.random.default_rng(100 + cfg)
    base = _randseq(rng, 3000)
    reads = []
    for _ in range(400):
        a = int(rng.integers(0, 2900))
        r = base[a:a + int(rng.integers(1, 150))]
        if rng.random() < 0.2:
            p = int(rng.integers(0, len(r)))
            r = r[:p] + "N" + r[p + 1:]
        reads.append(r)
    for canonical in (True, False):
        kmers, counts = oracle.count_reads(reads, K, canonical)
        rk, rc, kept, cut = ref.count_reads(cfg, reads, canonical, 3, n_workers=n_workers)
        assert np.array_equal(kmers, rk) and np.array_equal(counts, rc)
        okept, ocut = oracle.counter_to_set(kmers, counts, 3)
        assert np.array_equal(okept, kept) and ocut == cut


@pytest.mark.parametrize("cfg", [2, 3, 4, 5])
def test_sampled_and_full_sets(oracle, ref, cfg):
    K, N, _ = CONFIGS[cfg]
    rng = np.random.default_rng(200 + cfg)
    strs = [_randseq(rng, int(rng.integers(K, 2000))) for _ in range(60)]
    ids = rng.permutation(1 << N)[: (1 << N) // 50].astype(np.int32)
    for canonical in (True, False):
        ro, rk, size, weight = ref.sampled_set(cfg, strs, canonical, ids, n_workers=3)
        oo, ok = oracle.sampled_set(strs, K, N, canonical, ids)
        assert np.array_equal(ro, oo) and np.array_equal(rk, ok)
        rs, rh = ref.set_from_spss(cfg, strs, canonical, n_workers=2)
        os_ = oracle.set_from_spss(strs, K, canonical)
        assert np.array_equal(rs, os_) and rh == oracle.set_hash(os_)


@pytest.mark.parametrize("cfg", [2, 4])
def test_set_algebra(oracle, ref, cfg):
    K = CONFIGS[cfg][0]
    rng = np.random.default_rng(300 + cfg)
    pool = np.unique(rng.integers(0, 1 << (2 * K), 5000, dtype=np.uint64))
    a = np.sort(rng.choice(pool, 3000, replace=False))
    b = np.sort(rng.choice(pool, 3000, replace=False))
    assert np.array_equal(ref.set_op(cfg, "add", a, b, 2), oracle.set_add(a, b))
    assert np.array_equal(ref.set_op(cfg, "sub", a, b, 2), oracle.set_sub(a, b))
    assert np.array_equal(ref.set_op(cfg, "intersection", a, b, 2), oracle.set_intersection(a, b))
    assert ref.set_op(cfg, "diff", a, b) == oracle.set_diff(a, b)
    assert ref.set_op(cfg, "hash", a, b) == oracle.set_hash(a)


def test_reference_spss_is_valid(oracle, ref):
    """test/spss.cc:99-153 contract: the reference's SPSS spells each canonical k-mer once."""
    cfg = 2
    K = CONFIGS[cfg][0]
    rng = np.random.default_rng(5)
    kmers = oracle.set_from_spss([_randseq(rng, 4000)], K, True)
    spss, weight = ref.spss_from_set(cfg, kmers, True, n_workers=2)
    assert weight == sum(len(s) for s in spss)
    allk = oracle.spss_kmers(spss, K, True)
    assert len(allk) == len(kmers) and np.array_equal(np.sort(allk), kmers)


def test_repeated_bucket_ids_count_once(oracle, ref):
    """a bucket id listed twice fills one position only (kmer_set_compact.h:127-131), so the
    reference's position-wise merge (kmer_set_set.h:161-181) sees it once"""
    cfg = 2
    K, N, kb = CONFIGS[cfg]
    rng = np.random.default_rng(9)
    a = [_randseq(rng, 3000)]
    b = [a[0][:1500] + _randseq(rng, 1500)]
    ids = np.array([5, 7, 100, 5, 9000, 7], np.int32)
    ids = np.concatenate([ids, rng.permutation(1 << N)[:400].astype(np.int32)])
    ao, ak, _, _ = ref.sampled_set(cfg, a, True, ids)
    bo, bk, _, _ = ref.sampled_set(cfg, b, True, ids)
    want = sum(oracle.merge_count(ak[ao[i]:ao[i + 1]], bk[bo[i]:bo[i + 1]]) for i in range(len(ids)))
    ka, kb_ = oracle.set_from_spss(a, K, True), oracle.set_from_spss(b, K, True)
    oa, keysa = oracle.to_csr(ka, K, N, 2)
    ob, keysb = oracle.to_csr(kb_, K, N, 2)
    w, _ = oracle.pair_counts([oa, ob], [keysa, keysb], 2, 1 << N, bucket_ids=ids)
    assert w[0, 1] == want
