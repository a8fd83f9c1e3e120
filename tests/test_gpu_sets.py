"""P2 / P4 parity (GPU) through the C ABI against the oracle:
 * kmsc_set_from_spss  vs GetSampledKmerSet / GetKmerSetFromSPSS restatements
   (reference lib/core/kmer_set_compact.h:120-203, lib/core/spss.h:1861-1941)
 * kmsc_pair_split / kmsc_set_union / kmsc_set_diff vs KmerSet algebra
   (lib/core/kmer_set.h:164-219, 301-305; kmer_set_set.h:332-343)
plus the golden fixtures generated from the reference's own code."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
from _oracle import CONFIGS  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).parent / "golden" / "ref_golden.json").read_text())
KB = {5: 2, 9: 2, 15: 2, 19: 4, 23: 4, 31: 8}


@pytest.fixture(scope="module")
def ctx():
    import kmsc
    c = kmsc.Context(0)
    yield c
    c.close()


def _randseq(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


@pytest.mark.parametrize("K,N", [(15, 14), (19, 10), (23, 14), (31, 14), (9, 10), (5, 3)])
@pytest.mark.parametrize("canonical", [True, False])
def test_spss_to_set_dedup(ctx, oracle, K, N, canonical):
    rng = np.random.default_rng(K)
    strs = [_randseq(rng, int(rng.integers(K, 3000))) for _ in range(40)]
    strs += [strs[0], "A" * (K + 40), _randseq(rng, K)]  # duplicates, low complexity, exactly one k-mer
    s = ctx.set_from_spss(K, N, KB[K], strs, canonical=canonical, dedup=True)
    want = oracle.set_from_spss(strs, K, canonical)
    assert np.array_equal(s.to_kmers(), want)
    assert s.Size() == len(want) and s.Hash() == oracle.set_hash(want)


@pytest.mark.parametrize("K,N", [(15, 14), (23, 14), (19, 10)])
def test_spss_sampled_keeps_duplicates(ctx, oracle, K, N):
    """GetSampledKmerSet semantics: all buckets, duplicates preserved, each bucket sorted"""
    rng = np.random.default_rng(100 + K)
    strs = [_randseq(rng, int(rng.integers(K, 2000))) for _ in range(30)]
    strs += [strs[3], strs[3]]
    s = ctx.set_from_spss(K, N, KB[K], strs, canonical=True, dedup=False)
    ids = np.arange(1 << N, dtype=np.int32)
    offs, keys = oracle.sampled_set(strs, K, N, True, ids)
    go, gk = s.to_csr()
    assert np.array_equal(go, offs) and np.array_equal(gk.astype(np.uint64), keys)


def test_spss_bucket_range_shard(ctx, oracle):
    """a rank's prefix shard: only buckets in [lo, hi)"""
    K, N = 23, 14
    rng = np.random.default_rng(3)
    strs = [_randseq(rng, 20000)]
    full = oracle.set_from_spss(strs, K, True)
    lo, hi = 1000, 5000
    s = ctx.set_from_spss(K, N, 4, strs, bucket_lo=lo, bucket_hi=hi)
    b = full >> np.uint64(2 * K - N)
    assert np.array_equal(s.to_kmers(), full[(b >= lo) & (b < hi)])


def test_spss_edge_cases(ctx, oracle):
    import kmsc
    K, N = 15, 14
    assert ctx.set_from_spss(K, N, 2, []).Size() == 0
    assert ctx.set_from_spss(K, N, 2, ["ACGT"]).Size() == 0  # shorter than K
    assert ctx.set_from_spss(K, N, 2, ["A" * 15]).Size() == 1
    with pytest.raises(kmsc.KmscError):
        ctx.set_from_spss(K, N, 2, ["ACGTNACGTACGTACGTACGT"])


def test_spss_skewed_big_runs(ctx, oracle):
    """thousands of k-mers in one fine bucket: exercises the shared-memory and the
    global bitonic sort paths"""
    K, N = 23, 14
    rng = np.random.default_rng(8)
    prefix = "ACGTACGTACG"  # 11 fixed bases = 22 bits > N + 6
    strs = [prefix + _randseq(rng, 12) for _ in range(6000)]
    s = ctx.set_from_spss(K, N, 4, strs, canonical=False, dedup=True)
    want = oracle.set_from_spss(strs, K, False)
    assert np.array_equal(s.to_kmers(), want)
    s2 = ctx.set_from_spss(K, N, 4, strs[:3000], canonical=False, dedup=False)
    offs, keys = oracle.sampled_set(strs[:3000], K, N, False, np.arange(1 << N, dtype=np.int32))
    go, gk = s2.to_csr()
    assert np.array_equal(go, offs) and np.array_equal(gk.astype(np.uint64), keys)


def test_golden_compact(ctx):
    """fixtures produced by the reference's own GetSampledKmerSet / ToKmerSet"""
    for e in GOLD["compact"]:
        K, N, _ = CONFIGS[e["cfg"]]
        if e["bucket_ids"] != "reversed_all":
            continue
        strs = GOLD["compact_strings"][e["strings_id"]]
        s = ctx.set_from_spss(K, N, KB[K], strs, canonical=e["canonical"], dedup=False)
        offs, keys = s.to_csr()
        want_keys = np.array(e["keys"], np.uint64)
        pos = 0
        seen = 0
        for i, c in e["counts_nz"]:  # position i of the reversed id list = bucket 2^N - 1 - i
            b = (1 << N) - 1 - i
            assert offs[b + 1] - offs[b] == c
            assert np.array_equal(keys[offs[b]:offs[b + 1]].astype(np.uint64), want_keys[pos:pos + c])
            pos += c
            seen += c
        assert seen == s.Size()
        d = ctx.set_from_spss(K, N, KB[K], strs, canonical=e["canonical"], dedup=True)
        assert d.Size() == e["set_size"] and d.Hash() == e["set_hash"]


@pytest.mark.parametrize("K,N", [(15, 14), (23, 14), (31, 14), (19, 10), (5, 3)])
def test_pair_split_union_diff(ctx, oracle, K, N):
    import synth
    seqs = synth.phylogeny_sequences(3, 50000 if K > 5 else 300, p=0.02, seed=K)
    a, b, c = (synth.kmer_set_of(s, K) for s in seqs)
    da, db, dc = (ctx.set_from_kmers(K, N, KB[K], x) for x in (a, b, c))
    inter, am, bm = ctx.pair_split(da, db)
    wi = oracle.set_intersection(a, b)
    assert np.array_equal(inter.to_kmers(), wi)
    assert np.array_equal(am.to_kmers(), oracle.set_sub(a, wi))
    assert np.array_equal(bm.to_kmers(), oracle.set_sub(b, wi))
    assert inter.Hash() == oracle.set_hash(wi)
    assert ctx.set_diff(da, db) == oracle.set_diff(a, b)
    u = ctx.set_union([da, db, dc])
    assert np.array_equal(u.to_kmers(), oracle.set_add(oracle.set_add(a, b), c))
    # reconstruction identity of the greedy factoring (kmer_set_set.h:433-454)
    r = ctx.set_union([am, inter])
    assert np.array_equal(r.to_kmers(), a)
    # split outputs are valid inputs of P3
    w = ctx.pair_counts([inter, am, bm, da])
    assert w[0, 1] == 0 and w[0, 2] == 0 and w[1, 2] == 0 and w[0, 3] == len(wi)


def test_golden_setops(ctx):
    for e in GOLD["setops"]:
        K, N, _ = CONFIGS[e["cfg"]]
        a, b = np.array(e["a"], np.uint64), np.array(e["b"], np.uint64)
        da, db = ctx.set_from_kmers(K, N, KB[K], a), ctx.set_from_kmers(K, N, KB[K], b)
        inter, am, bm = ctx.pair_split(da, db)
        assert inter.to_kmers().tolist() == e["intersection"]
        assert am.to_kmers().tolist() == e["sub"]
        assert ctx.set_union([da, db]).to_kmers().tolist() == e["add"]
        assert ctx.set_diff(da, db) == e["diff"] and da.Hash() == e["hash"]


def test_split_empty_and_disjoint(ctx, oracle):
    K, N = 23, 14
    a = np.array([1, 2, 3, (7 << 32) | 5], np.uint64)
    e = np.zeros(0, np.uint64)
    da, de = ctx.set_from_kmers(K, N, 4, a), ctx.set_from_kmers(K, N, 4, e)
    i, x, y = ctx.pair_split(da, de)
    assert i.Size() == 0 and np.array_equal(x.to_kmers(), a) and y.Size() == 0
    i, x, y = ctx.pair_split(da, da)
    assert np.array_equal(i.to_kmers(), a) and x.Size() == 0 and y.Size() == 0


def _check_split(ctx, oracle, K, N, pairs_km, hint, want=(True, True, True)):
    dj = [ctx.set_from_kmers(K, N, KB[K], a) for a, _ in pairs_km]
    dk = [ctx.set_from_kmers(K, N, KB[K], b) for _, b in pairs_km]
    inter, jm, km = ctx.pair_split_batch(dj, dk, inter_hint=hint, want_inter=want[0], want_j=want[1], want_k=want[2])
    for p, (a, b) in enumerate(pairs_km):
        wi = oracle.set_intersection(a, b)
        if want[0]:
            assert np.array_equal(inter[p].to_kmers(), wi), f"pair {p}: intersection"
            assert inter[p].Hash() == oracle.set_hash(wi)
        if want[1]:
            assert np.array_equal(jm[p].to_kmers(), oracle.set_sub(a, wi)), f"pair {p}: j minus"
        if want[2]:
            assert np.array_equal(km[p].to_kmers(), oracle.set_sub(b, wi)), f"pair {p}: k minus"
    # the offset levels of the new sets are exact: P3 over them gives the set algebra back
    if all(want) and len(pairs_km) >= 1:
        w = ctx.pair_counts([inter[0], jm[0], km[0], dj[0], dk[0]])
        n = len(oracle.set_intersection(*pairs_km[0]))
        assert w[0, 1] == 0 and w[0, 2] == 0 and w[1, 2] == 0 and w[0, 3] == n and w[0, 4] == n
        assert w[1, 3] == len(pairs_km[0][0]) - n and w[2, 4] == len(pairs_km[0][1]) - n


@pytest.fixture(params=["auto", "0", "2", "5"])
def qlog(request, monkeypatch):
    """threads per fine bucket of the split kernel (2^qlog): chosen from the density, or forced"""
    if request.param == "auto":
        monkeypatch.delenv("KMSC_SPLIT_QLOG", raising=False)
    else:
        monkeypatch.setenv("KMSC_SPLIT_QLOG", request.param)
    return request.param


@pytest.mark.parametrize("K,N", [(15, 14), (23, 14), (31, 14), (19, 10)])
def test_pair_split_batch(ctx, oracle, K, N, qlog):
    """kmsc_pair_split_batch: several pairs per pass, with and without |j & k| hints (right and
    wrong), any subset of outputs; reference semantics kmer_set_set.h:332-343"""
    import synth
    seqs = synth.phylogeny_sequences(6, 60000, p=0.01, seed=100 + K)
    ks = [synth.kmer_set_of(s, K) for s in seqs]
    e = np.zeros(0, np.uint64)
    pairs = [(ks[0], ks[1]), (ks[2], ks[0]), (ks[3], ks[3]), (ks[4], e), (e, ks[5]), (e, e), (ks[1], ks[5][::3].copy())]
    exact = np.array([len(oracle.set_intersection(a, b)) for a, b in pairs], np.int64)
    _check_split(ctx, oracle, K, N, pairs, None)
    _check_split(ctx, oracle, K, N, pairs, exact)
    wrong = exact.copy()
    wrong[0] = max(0, wrong[0] - 7); wrong[2] += 5; wrong[6] = 1 << 40; wrong[1] = -1
    _check_split(ctx, oracle, K, N, pairs, wrong)
    _check_split(ctx, oracle, K, N, pairs[:3], exact[:3], want=(False, True, True))
    _check_split(ctx, oracle, K, N, pairs[:3], None, want=(True, False, False))


def test_pair_split_on_a_prefix_shard(ctx, oracle, qlog):
    """sets restricted to a bucket range (one rank's shard of a multi-GPU job): all keys sit in a
    fraction of the fine buckets, the kernel puts several threads on each"""
    import synth
    K, N = 23, 14
    seqs = synth.phylogeny_sequences(4, 400000, p=0.01, seed=77)
    lo, hi = 700, 1100
    dev, kms = [], []
    for s in seqs:
        dev.append(ctx.set_from_spss(K, N, 4, [synth.to_ascii(s).decode()], bucket_lo=lo, bucket_hi=hi))
        km = synth.kmer_set_of(s, K)
        b = (km >> np.uint64(2 * K - N)).astype(np.int64)
        kms.append(km[(b >= lo) & (b < hi)])
    for d, km in zip(dev, kms):
        assert np.array_equal(d.to_kmers(), km)
    js, ks = [dev[0], dev[1], dev[3]], [dev[1], dev[2], dev[0]]
    inter, jm, km_ = ctx.pair_split_batch(js, ks)
    for p, (a, b) in enumerate([(0, 1), (1, 2), (3, 0)]):
        wi = oracle.set_intersection(kms[a], kms[b])
        assert np.array_equal(inter[p].to_kmers(), wi)
        assert np.array_equal(jm[p].to_kmers(), oracle.set_sub(kms[a], wi))
        assert np.array_equal(km_[p].to_kmers(), oracle.set_sub(kms[b], wi))
    w = ctx.pair_counts([inter[0], jm[0], km_[0], dev[0], dev[1]])
    assert w[0, 1] == 0 and w[0, 3] == len(oracle.set_intersection(kms[0], kms[1])) and w[1, 4] == 0


def test_pair_split_dense_runs(ctx, oracle, qlog):
    """keys concentrated in a few fine buckets: chunks that need several rounds, and single fine
    buckets larger than the shared-memory tile (one-thread path)"""
    K, N = 23, 14
    rng = np.random.default_rng(5)

    def cluster(prefix20, n):  # n distinct 46-bit values sharing their top 20 bits = one finest fine bucket
        low = rng.choice(1 << 26, size=n, replace=False).astype(np.uint64)
        return (np.uint64(prefix20) << np.uint64(26)) | low

    base = np.concatenate([cluster(5, 3000), cluster(6, 3500), cluster(7, 2500), cluster(9, 12000), cluster(4000, 30000),
                           rng.choice(1 << 46, size=20000, replace=False).astype(np.uint64)])
    base = np.unique(base)
    a = base[rng.random(len(base)) < 0.9]
    b = base[rng.random(len(base)) < 0.8]
    c = np.unique(np.concatenate([cluster(9, 9000), cluster(4000, 100), a[::7]]))
    pairs = [(a, b), (b, c), (c, a)]
    _check_split(ctx, oracle, K, N, pairs, None)
    exact = np.array([len(oracle.set_intersection(x, y)) for x, y in pairs], np.int64)
    _check_split(ctx, oracle, K, N, pairs, exact)


def test_pair_split_full_size_properties(ctx):
    """BASELINE config 2 sizes (10 M canonical 23-mers per set): size-independent properties --
    |n| = W[j][k] from P3, |j\\n| = |j| - |n|, hash(n) ^ hash(j\\n) = hash(j), outputs disjoint"""
    import synth
    K, N = 23, 14
    seqs = synth.phylogeny_sequences(3, 10_000_000 + K - 1, p=0.002, seed=9)
    sets = [ctx.set_from_spss(K, N, 4, [synth.to_ascii(s).decode()]) for s in seqs]
    W = ctx.pair_counts(sets)
    js, ks = [sets[0], sets[0], sets[1]], [sets[1], sets[2], sets[2]]
    hint = np.array([W[0, 1], W[0, 2], W[1, 2]], np.int64)
    for h in (hint, None):
        inter, jm, km = ctx.pair_split_batch(js, ks, inter_hint=h)
        for p in range(3):
            assert inter[p].Size() == hint[p]
            assert jm[p].Size() == js[p].Size() - hint[p] and km[p].Size() == ks[p].Size() - hint[p]
            assert inter[p].Hash() ^ jm[p].Hash() == js[p].Hash()
            assert inter[p].Hash() ^ km[p].Hash() == ks[p].Hash()
            w = ctx.pair_counts([inter[p], jm[p], km[p]])
            assert w[0, 1] == 0 and w[0, 2] == 0 and w[1, 2] == 0
            for s in (inter[p], jm[p], km[p]):
                s.free()


def test_export_import_bucket_ranges(ctx, oracle):
    """multi-GPU exchange helpers: a set cut into bucket ranges, every range exported to device
    buffers and imported as a restricted set; the ranges partition the set exactly"""
    import synth
    import torch
    K, N, kb = 23, 14, 4
    km = synth.kmer_set_of(synth.random_genome(50000, K), K)
    offs, keys = synth.csr_of(km, K, N, kb)
    s = ctx.set_from_csr(K, N, kb, offs, keys)
    cuts = np.array([0, 100, 100, 9000, 1 << N], np.int32)   # one empty range
    ko = ctx.set_bucket_offsets(s, cuts)
    assert np.array_equal(ko, offs[cuts])
    dev = torch.device("cuda", 0)
    parts = []
    for q in range(len(cuts) - 1):
        lo, hi = int(cuts[q]), int(cuts[q + 1])
        cnt = int(ko[q + 1] - ko[q])
        d_offs = torch.empty(hi - lo + 1, dtype=torch.int32, device=dev)
        d_keys = torch.empty(max(1, cnt), dtype=torch.int32, device=dev)
        ctx.set_export_range(s, lo, hi, int(ko[q]), int(ko[q + 1]), d_offs.data_ptr(), d_keys.data_ptr())
        ctx.sync()
        torch.cuda.synchronize()
        p = ctx.set_import_range(K, N, kb, lo, hi, d_offs.data_ptr(), d_keys.data_ptr(), cnt)
        po, pk = p.to_csr()
        want_o = np.clip(offs, ko[q], ko[q + 1]) - ko[q]
        assert np.array_equal(po, want_o) and np.array_equal(pk, keys[ko[q]:ko[q + 1]])
        parts.append(p)
    u = ctx.set_union(parts)
    assert np.array_equal(u.to_kmers(), km)
    # intersection counts add up over the ranges (what the all-reduce relies on)
    total = sum(int(ctx.pair_counts([p, s])[0, 1]) for p in parts)
    assert total == len(km)


@pytest.mark.parametrize("K,N,canonical", [(15, 14, True), (23, 14, True), (31, 14, True), (9, 10, False), (5, 3, True)])
def test_neighbor_table(ctx, oracle, K, N, canonical):
    """kmsc_set_neighbors vs Kmer::Next / Prev / Canonical (reference lib/core/kmer.h:133-186) + set lookup"""
    import synth
    km = synth.kmer_set_of(synth.random_genome(3000, K), K, canonical=canonical)
    offs, keys = synth.csr_of(km, K, N, KB[K])
    s = ctx.set_from_csr(K, N, KB[K], offs, keys)
    nb = ctx.set_neighbors(s, canonical=canonical)
    index = {int(v): i for i, v in enumerate(km)}
    rng = np.random.default_rng(K)
    for i in rng.integers(0, len(km), 300):
        v = int(km[i])
        for c in range(4):
            for d, w in ((0, oracle.next(v, K, "ACGT"[c])), (4, oracle.prev(v, K, "ACGT"[c]))):
                q, flip = w, 0
                if canonical:
                    q = oracle.canonical(w, K)
                    flip = int(q != w)
                want = (index[q] << 1 | flip) if q in index else -1
                assert nb[i, d + c] == want, (i, d, c)


def test_set_from_csr_rejects_bad_input(ctx):
    """unsorted keys, keys wider than 2K-N bits and impossible shapes are refused instead of
    producing wrong matrices later; K = 32 with N = 0 (64 key bits) is not supported"""
    import kmsc
    K, N = 23, 14
    offs = np.zeros((1 << N) + 1, np.int64)
    offs[6:] = 2
    with pytest.raises(kmsc.KmscError):
        ctx.set_from_csr(K, N, 4, offs, np.array([9, 3], np.uint32))           # descending inside bucket 5
    assert ctx.set_from_csr(K, N, 4, offs, np.array([3, 9], np.uint32)).Size() == 2
    with pytest.raises(kmsc.KmscError):
        ctx.set_from_csr(15, 14, 4, offs, np.array([3, 1 << 20], np.uint32))  # 2K-N = 16 bits
    with pytest.raises(kmsc.KmscError):
        ctx.set_from_csr(32, 0, 8, np.array([0, 1], np.int64), np.array([5], np.uint64))
    with pytest.raises(kmsc.KmscError):
        ctx.set_from_csr(23, 40, 4, offs, np.array([3, 9], np.uint32))


def test_split_and_union_refuse_multisets(ctx):
    """split / diff / union assume true sets; a set holding a key twice is refused"""
    import kmsc
    K, N = 23, 14
    offs = np.zeros((1 << N) + 1, np.int64)
    offs[1:] = 3
    multi = ctx.set_from_csr(K, N, 4, offs, np.array([4, 4, 9], np.uint32))
    plain = ctx.set_from_csr(K, N, 4, offs, np.array([4, 5, 9], np.uint32))
    for call in (lambda: ctx.pair_split(multi, plain), lambda: ctx.pair_split(plain, multi),
                 lambda: ctx.set_diff(multi, plain), lambda: ctx.set_union([plain, multi])):
        with pytest.raises(kmsc.KmscError):
            call()
    assert ctx.set_union([plain, plain]).Size() == 3


@pytest.mark.parametrize("m", [3, 8, 16, 17, 40])
def test_multiway_union(ctx, oracle, m):
    """kmsc_set_union of m >= 3 sets: one pass pair per group of <= 16 (thread per finest bucket merging the m
    runs) instead of a left fold; equal to the numpy union, incl. empty and identical inputs"""
    import synth
    K, N = 23, 14
    rng = np.random.default_rng(m)
    seqs = synth.phylogeny_sequences(m, 20000, p=0.02, seed=100 + m)
    kms = [synth.kmer_set_of(s, K) for s in seqs]
    kms[1] = np.zeros(0, np.uint64)
    kms[2] = kms[0].copy()
    dev = []
    for km in kms:
        offs, keys = synth.csr_of(km, K, N, 4)
        dev.append(ctx.set_from_csr(K, N, 4, offs, keys))
    u = ctx.set_union(dev)
    want = np.unique(np.concatenate(kms))
    assert np.array_equal(u.to_kmers(), want)
    assert u.Hash() == oracle.set_hash(want)
    # the result is a full set: it feeds the other kernels
    assert int(ctx.pair_counts([u, dev[0]])[0, 1]) == len(kms[0])
