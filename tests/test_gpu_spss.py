"""f1 parity (GPU): SPSS construction on the device (kmsc_spss_build / kmsc_spss_fetch) against the property
the reference's own tests check for GetSPSS / GetSPSSCanonical (reference test/spss.cc:57-68, 113-124): the
strings spell every k-mer of the set exactly once. The decode side is the oracle's restatement of
GetKmerSetFromSPSS (lib/core/spss.h:1861-1941), so the check does not go through the product's own decoder."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))

pytestmark = pytest.mark.gpu

KB = {4: 2, 5: 2, 9: 2, 15: 2, 19: 4, 23: 4, 31: 8}
N_OF = {4: 3, 5: 3, 9: 10, 15: 14, 19: 10, 23: 14, 31: 14}


@pytest.fixture(scope="module")
def ctx():
    import kmsc
    c = kmsc.Context(0)
    yield c
    c.close()


def _check_spss(ctx, oracle, km, K, canonical, rounds=0):
    import synth
    N = N_OF[K]
    offs, keys = synth.csr_of(km, K, N, KB[K])
    s = ctx.set_from_csr(K, N, KB[K], offs, keys)
    strs = ctx.spss_build(s, canonical=canonical, rounds=rounds)
    s.free()
    assert all(len(x) >= K for x in strs)
    # every k-mer exactly once: as many k-mer positions as k-mers, and the decoded set is the set
    assert sum(len(x) - K + 1 for x in strs) == len(km)
    got = oracle.set_from_spss(strs, K, canonical)
    assert np.array_equal(got, km)
    return strs


@pytest.mark.parametrize("K,canonical", [(15, True), (23, True), (31, True), (9, False), (5, True), (19, False), (4, True)])
@pytest.mark.parametrize("rounds", [1, 0])
def test_spss_spells_every_kmer_once(ctx, oracle, K, canonical, rounds):
    import synth
    seqs = synth.phylogeny_sequences(3, 20000, p=0.01, seed=K)
    km = np.unique(np.concatenate([synth.kmers_of(s, K, canonical) for s in seqs]))
    strs = _check_spss(ctx, oracle, km, K, canonical, rounds)
    if K >= 15:
        # a mutated genome compacts: far fewer strings than k-mers
        assert len(strs) * 10 < len(km)


def test_spss_unitig_round_vs_more_rounds(ctx, oracle):
    """later matching rounds only ever join strings: never more strings than the unitig round alone"""
    import synth
    K = 23
    seqs = synth.phylogeny_sequences(4, 50000, p=0.01, seed=3)
    km = np.unique(np.concatenate([synth.kmers_of(s, K, True) for s in seqs]))
    u = _check_spss(ctx, oracle, km, K, True, rounds=1)
    m = _check_spss(ctx, oracle, km, K, True, rounds=8)
    assert len(m) <= len(u)
    assert sum(map(len, m)) <= sum(map(len, u))


@pytest.mark.parametrize("canonical", [True, False])
def test_spss_cycles_and_self_loops(ctx, oracle, canonical):
    """circular sequences (every port linked: pure cycles), homopolymers (a k-mer that follows itself),
    tandem repeats and a palindrome-rich text"""
    import synth
    K = 9
    rng = np.random.default_rng(1)
    parts = []
    for L in (40, 64, 100):           # circular: the k-mers of S + S[:K-1] form one cycle
        c = rng.integers(0, 4, L, dtype=np.uint8)
        parts.append(np.concatenate([c, c[: K - 1]]))
    parts.append(np.zeros(30, np.uint8))                         # AAAA...
    parts.append(np.full(30, 3, np.uint8))                       # TTTT... (the same canonical k-mer)
    parts.append(np.tile(np.array([0, 1], np.uint8), 20))        # ACACAC...
    parts.append(np.tile(np.array([0, 3], np.uint8), 20))        # ATATAT... (its own reverse complement)
    parts.append(np.tile(np.array([0, 1, 2], np.uint8), 15))     # ACGACG...
    km = np.unique(np.concatenate([synth.kmers_of(p, K, canonical) for p in parts]))
    _check_spss(ctx, oracle, km, K, canonical, rounds=0)
    _check_spss(ctx, oracle, km, K, canonical, rounds=1)


def test_spss_even_k_palindromes(ctx, oracle):
    """even K: k-mers equal to their own reverse complement have two ports reaching the same neighbours"""
    import synth
    K = 4
    km = np.arange(256, dtype=np.uint64)   # every 4-mer
    canon = np.unique(np.array([min(int(v), int(oracle.canonical(int(v), K))) for v in km], dtype=np.uint64))
    _check_spss(ctx, oracle, canon, K, True)
    _check_spss(ctx, oracle, km, K, False)


def test_spss_empty_and_single(ctx, oracle):
    import synth
    K, N = 23, 14
    s = ctx.set_from_kmers(K, N, 4, np.zeros(0, np.uint64))
    assert ctx.spss_build(s) == []
    s.free()
    km = np.array([12345678901], np.uint64)
    strs = _check_spss(ctx, oracle, km, K, False)
    assert len(strs) == 1 and len(strs[0]) == K


def test_spss_large_and_deterministic(ctx, oracle):
    """2 M k-mers: long paths (deep pointer jumping), same output twice, decoded by the product's own P2 as well"""
    import synth
    K, N = 23, 14
    seqs = synth.phylogeny_sequences(2, 1_000_000, p=0.002, seed=9)
    km = np.unique(np.concatenate([synth.kmers_of(s, K, True) for s in seqs]))
    offs, keys = synth.csr_of(km, K, N, 4)
    s = ctx.set_from_csr(K, N, 4, offs, keys)
    a = ctx.spss_build(s, canonical=True)
    b = ctx.spss_build(s, canonical=True)
    assert a == b
    assert sum(len(x) - K + 1 for x in a) == len(km)
    back = ctx.set_from_spss(K, N, 4, a, canonical=True, dedup=True)
    assert np.array_equal(back.to_kmers(), km)
    assert len(a) * 50 < len(km)
    back.free()
    s.free()


def test_spss_size_vs_reference(ctx, oracle, ref):
    """the unmodified reference's GetSPSSCanonical (fast = unitigs + greedy joins, lib/core/spss.h:1039-1858)
    on the same set: both outputs are valid; ours may not be much larger (Weight = characters decides the
    greedy driver's stop rule, lib/core/kmer_set_set.h:332-343)"""
    import synth
    from _oracle import CONFIGS
    cfg = next(c for c, v in CONFIGS.items() if v[0] == 15 and v[1] == 14)
    K = 15
    seqs = synth.phylogeny_sequences(3, 60000, p=0.01, seed=21)
    km = np.unique(np.concatenate([synth.kmers_of(s, K, True) for s in seqs]))
    ours = _check_spss(ctx, oracle, km, K, True)
    theirs, _w = ref.spss_from_set(cfg, km, True, fast=True)
    assert np.array_equal(oracle.set_from_spss(theirs, K, True), km)
    ours_chars, theirs_chars = sum(map(len, ours)), sum(map(len, theirs))
    print(f"SPSS of {len(km)} 15-mers: ours {len(ours)} strings / {ours_chars} chars, reference {len(theirs)} / {theirs_chars}")
    assert ours_chars <= 1.15 * theirs_chars


@pytest.mark.parametrize("K,canonical", [(23, True), (9, False), (31, True)])
def test_spss_fetch_packed(ctx, oracle, K, canonical):
    """kmsc_spss_fetch_packed = the strings of kmsc_spss_fetch packed 2 bits per base, 32 per word, first base in
    the top bits, strings back to back (KmerSetCompact's container), and decodes to the set through kmsc_set_from_packed"""
    import synth
    from test_gpu_decode_batch import pack
    N = N_OF[K]
    km = np.unique(np.concatenate([synth.kmers_of(s, K, canonical) for s in synth.phylogeny_sequences(3, 30000, p=0.01, seed=K + 1)]))
    offs, keys = synth.csr_of(km, K, N, KB[K])
    s = ctx.set_from_csr(K, N, KB[K], offs, keys)
    strs = ctx.spss_build(s, canonical=canonical)
    words, str_offs = ctx.spss_build_packed(s, canonical=canonical)
    want_words, want_offs = pack(strs)
    nw = (int(str_offs[-1]) + 31) // 32
    assert np.array_equal(str_offs, want_offs) and np.array_equal(words[:nw], want_words[:nw])
    back = ctx.set_from_packed(K, N, KB[K], words, str_offs, canonical=canonical)
    assert np.array_equal(back.to_kmers(), km)
    back.free()
    s.free()


def test_spss_refuses_multisets(ctx):
    """a device set built with dedup = 0 from a text that repeats a k-mer is a multiset: no SPSS of it"""
    import kmsc
    K, N = 23, 14
    s = ctx.set_from_spss(K, N, 4, ["ACGTACGTTGCATGCAAGCTTGCATTT", "ACGTACGTTGCATGCAAGCTTGCATTT"], canonical=True, dedup=False)
    with pytest.raises(kmsc.KmscError):
        ctx.spss_build(s)
    s.free()
