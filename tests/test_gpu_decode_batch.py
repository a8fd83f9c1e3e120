"""P2 parity (GPU) of the packed / batched decode entry points against the oracle:
kmsc_set_from_packed and kmsc_sets_from_packed_batch vs the restatements of
KmerSetCompact::GetSampledKmerSet (reference lib/core/kmer_set_compact.h:120-203; dedup = 0)
and GetKmerSetFromSPSS (lib/core/spss.h:1861-1941; dedup = 1). Both the staged partition
sort and the general pipeline it falls back to (KMSC_P2_LEGACY=1) are run."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))

pytestmark = pytest.mark.gpu
KB = {5: 2, 9: 2, 15: 2, 19: 4, 23: 4, 31: 8}
CODE = {c: i for i, c in enumerate("ACGT")}


@pytest.fixture(scope="module")
def ctx():
    import kmsc
    c = kmsc.Context(0)
    yield c
    c.close()


@pytest.fixture(params=["partition", "legacy"])
def path(request, monkeypatch):
    if request.param == "legacy":
        monkeypatch.setenv("KMSC_P2_LEGACY", "1")
    return request.param


def pack(strs):
    """KmerSetCompact's in-memory form: 2 bits per base, 32 bases per uint64, first base on top"""
    text = "".join(strs)
    codes = np.array([CODE[c] for c in text], np.uint64) if text else np.zeros(0, np.uint64)
    n = len(codes)
    nw = (n + 31) // 32
    pad = np.zeros(nw * 32, np.uint64)
    pad[:n] = codes
    shifts = np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64)
    words = (pad.reshape(nw, 32) << shifts).sum(axis=1, dtype=np.uint64) if nw else np.zeros(0, np.uint64)
    words = np.concatenate([words, np.zeros(2, np.uint64)])
    offs = np.zeros(len(strs) + 1, np.int64)
    np.cumsum([len(s) for s in strs], out=offs[1:])
    return words, offs


def randseq(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


@pytest.mark.parametrize("K,N", [(15, 14), (19, 10), (23, 14), (31, 14), (9, 10), (5, 3)])
@pytest.mark.parametrize("canonical", [True, False])
def test_set_from_packed_vs_oracle(ctx, oracle, path, K, N, canonical):
    rng = np.random.default_rng(7 * K + canonical)
    strs = [randseq(rng, int(rng.integers(1, 4000))) for _ in range(30)]   # some shorter than K
    strs += [strs[1], "", "C" * (K + 70), randseq(rng, K)]
    words, offs = pack(strs)
    ids = np.arange(1 << N, dtype=np.int32)
    want_offs, want_keys = oracle.sampled_set(strs, K, N, canonical, ids)
    s = ctx.set_from_packed(K, N, KB[K], words, offs, canonical=canonical, dedup=False)
    go, gk = s.to_csr()
    assert np.array_equal(go, want_offs) and np.array_equal(gk.astype(np.uint64), want_keys)
    d = ctx.set_from_packed(K, N, KB[K], words, offs, canonical=canonical, dedup=True)
    want = oracle.set_from_spss(strs, K, canonical)
    assert np.array_equal(d.to_kmers(), want) and d.Hash() == oracle.set_hash(want)


@pytest.mark.parametrize("K,N", [(23, 14), (15, 14), (31, 14), (19, 10)])
@pytest.mark.parametrize("dedup", [True, False])
def test_batch_vs_oracle(ctx, oracle, path, K, N, dedup):
    """sets of very different sizes in one batch, an empty one, one with repeated k-mers"""
    rng = np.random.default_rng(31 * K + dedup)
    jobs = []
    for j in range(11):
        n_str = int(rng.integers(1, 12))
        jobs.append([randseq(rng, int(rng.integers(K, 6000))) for _ in range(n_str)])
    jobs[3] = []
    jobs[5] = jobs[5] + jobs[5][:2] + ["G" * (K + 300)]      # repeats: the dedup fallback of one job
    jobs[7] = [randseq(rng, 120000)]
    packed = [pack(s) for s in jobs]
    sets = ctx.sets_from_packed_batch(K, N, KB[K], [p[0] for p in packed], [p[1] for p in packed], dedup=dedup)
    ids = np.arange(1 << N, dtype=np.int32)
    for strs, s in zip(jobs, sets):
        if dedup:
            want = oracle.set_from_spss(strs, K, True)
            assert np.array_equal(s.to_kmers(), want)
            assert s.Hash() == oracle.set_hash(want)
        else:
            wo, wk = oracle.sampled_set(strs, K, N, True, ids)
            go, gk = s.to_csr()
            assert np.array_equal(go, wo) and np.array_equal(gk.astype(np.uint64), wk)


def test_batch_bucket_range_and_pair_counts(ctx, oracle, path):
    """a rank's prefix shard out of the batch call feeds P3 like sets built one by one"""
    K, N = 23, 14
    rng = np.random.default_rng(5)
    base = randseq(rng, 60000)
    jobs = []
    for j in range(6):
        b = list(base)
        for p in rng.integers(0, len(b), 200):
            b[p] = "ACGT"[(CODE[b[p]] + 1 + j) & 3]
        jobs.append(["".join(b)])
    packed = [pack(s) for s in jobs]
    lo, hi = 3000, 11000
    sets = ctx.sets_from_packed_batch(K, N, 4, [p[0] for p in packed], [p[1] for p in packed], bucket_lo=lo, bucket_hi=hi)
    offs_l, keys_l = [], []
    for strs, s in zip(jobs, sets):
        full = oracle.set_from_spss(strs, K, True)
        b = full >> np.uint64(2 * K - N)
        want = full[(b >= lo) & (b < hi)]
        assert np.array_equal(s.to_kmers(), want)
        o, k = oracle.to_csr(want, K, N, 4)
        offs_l.append(o)
        keys_l.append(k)
    want_w, _ = oracle.pair_counts(offs_l, keys_l, 4, 1 << N)
    got = ctx.pair_counts(sets)
    iu = np.triu_indices(len(sets), 1)
    assert np.array_equal(got[iu], want_w[iu])


def test_batch_two_million_kmers(ctx, oracle):
    """hundreds of tiles and every first-level bin in use: 2 sets x 2 M canonical 23-mers vs the oracle"""
    import synth
    K, N = 23, 14
    seqs = synth.phylogeny_sequences(2, 2_000_000 + K - 1, 0.002)
    jobs = [[synth.to_ascii(s).decode()] for s in seqs]
    packed = [pack(s) for s in jobs]
    sets = ctx.sets_from_packed_batch(K, N, 4, [p[0] for p in packed], [p[1] for p in packed])
    for codes, s in zip(seqs, sets):
        want = synth.kmer_set_of(codes, K, True)
        if len(want) != 2_000_000:   # a repeated k-mer: dedup must agree with the oracle anyway
            pass
        assert np.array_equal(s.to_kmers(), want)
    one = ctx.set_from_packed(K, N, 4, packed[0][0], packed[0][1])
    assert np.array_equal(one.to_kmers(), synth.kmer_set_of(seqs[0], K, True))


def test_batch_low_complexity_long_runs(ctx, oracle, path):
    """one fine bucket holding thousands of distinct keys and a k-mer repeated thousands of times:
    the CTA-wide sort of long sub-bin runs and the heap-sort overflow"""
    K, N = 23, 14
    rng = np.random.default_rng(9)
    prefix = "ACGTACGTACGTAC"  # 14 fixed bases = 28 bits: one fine bucket, few sub-bins
    strs = [prefix + randseq(rng, 9) for _ in range(5000)] + ["T" * 3000] + [prefix + "AAAAAAAAA"] * 700
    words, offs = pack(strs)
    s = ctx.sets_from_packed_batch(K, N, 4, [words], [offs], canonical=False, dedup=False)[0]
    wo, wk = oracle.sampled_set(strs, K, N, False, np.arange(1 << N, dtype=np.int32))
    go, gk = s.to_csr()
    assert np.array_equal(go, wo) and np.array_equal(gk.astype(np.uint64), wk)
    d = ctx.sets_from_packed_batch(K, N, 4, [words], [offs], canonical=False, dedup=True)[0]
    assert np.array_equal(d.to_kmers(), oracle.set_from_spss(strs, K, False))

