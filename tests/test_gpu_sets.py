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
