"""P5 / P6 parity (GPU) through the C ABI.
 * kmsc_bitmap_gram (k <= 15 dense bitmaps, AND + POPC Gram) must equal the exact all-bucket
   intersection matrix: the oracle's restatement of GetEdgeWeight over all buckets
   (reference lib/core/kmer_set_set.h:158-219) and kmsc_pair_counts.
 * kmsc_codec_encode / kmsc_codec_decode (delta + streamvbyte-style byte codes; container in
   csrc/codec.cu) must match the oracle's CPU restatement byte for byte and round-trip."""
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


def _sets(ctx, kmer_sets, K, N, kb):
    import synth
    dev, offs_l, keys_l = [], [], []
    for km in kmer_sets:
        offs, keys = synth.csr_of(km, K, N, kb)
        dev.append(ctx.set_from_csr(K, N, kb, offs, keys))
        offs_l.append(offs)
        keys_l.append(keys)
    return dev, offs_l, keys_l


@pytest.mark.parametrize("K,N,kb,n_sets", [(15, 14, 2, 8), (11, 8, 2, 70), (9, 10, 2, 5), (5, 3, 2, 3), (9, 10, 2, 300), (11, 8, 2, 129)])
def test_bitmap_gram_matches_merge(ctx, oracle, K, N, kb, n_sets):
    import synth
    glen = 40000 if n_sets <= 70 else 6000   # many sets: keep the CPU oracle fast
    seqs = synth.window_sequences(n_sets, glen, glen // 2, p=0.01, seed=K)
    sets = [synth.kmer_set_of(s, K) for s in seqs]
    sets[-1] = sets[-1][:0]  # an empty set
    dev, offs_l, keys_l = _sets(ctx, sets, K, N, kb)
    got = ctx.bitmap_gram(dev)
    want, _ = oracle.pair_counts(offs_l, keys_l, kb, 1 << N, n_threads=8)
    iu = np.triu_indices(n_sets, 1)
    assert np.array_equal(got[iu], want[iu])
    assert np.array_equal(got, got.T)
    assert [int(got[i, i]) for i in range(n_sets)] == [len(s) for s in sets]
    assert np.array_equal(got, ctx.pair_counts(dev))
    for d in dev:
        d.free()


def test_bitmap_gram_rejects_large_k(ctx):
    import kmsc
    import synth
    km = synth.kmer_set_of(synth.random_genome(2000, 23), 23)
    offs, keys = synth.csr_of(km, 23, 14, 4)
    s = ctx.set_from_csr(23, 14, 4, offs, keys)
    with pytest.raises(kmsc.KmscError):
        ctx.bitmap_gram([s, s])
    s.free()


@pytest.mark.parametrize("K,N,kb", [(15, 14, 2), (23, 14, 4), (31, 14, 8), (19, 10, 4), (5, 3, 2)])
def test_codec_bytes_and_roundtrip(ctx, oracle, K, N, kb):
    import synth
    km = synth.kmer_set_of(synth.random_genome(60000, K), K)
    offs, keys = synth.csr_of(km, K, N, kb)
    s = ctx.set_from_csr(K, N, kb, offs, keys)
    data = ctx.codec_encode(s)
    assert data == oracle.codec_encode(K, N, kb, offs, keys)
    back = ctx.codec_decode(data)
    o2, k2 = back.to_csr()
    assert np.array_equal(o2, offs) and np.array_equal(k2, keys)
    assert back.Size() == len(km) and back.Hash() == oracle.set_hash(km)
    # the decoded set is a full citizen: usable by the other primitives
    inter, a, b = ctx.pair_split(s, back)
    assert inter.Size() == len(km) and a.Size() == 0 and b.Size() == 0


def test_codec_empty_and_wide_deltas(ctx, oracle):
    K, N, kb = 23, 14, 4
    empty = np.zeros(0, np.uint64)
    import synth
    for km in (empty, np.array([0, 1, 0xFFFFFFFF, (3 << 32) | 5, (3 << 32) | 0x01000005, (1 << 46) - 1], np.uint64)):
        offs, keys = synth.csr_of(km, K, N, kb)
        s = ctx.set_from_csr(K, N, kb, offs, keys)
        data = ctx.codec_encode(s)
        assert data == oracle.codec_encode(K, N, kb, offs, keys)
        back = ctx.codec_decode(data)
        assert np.array_equal(back.to_kmers(), km)


def test_codec_rejects_corrupt_input(ctx, oracle):
    import kmsc
    import synth
    km = synth.kmer_set_of(synth.random_genome(3000, 15), 15)
    offs, keys = synth.csr_of(km, 15, 14, 2)
    data = oracle.codec_encode(15, 14, 2, offs, keys)
    with pytest.raises(kmsc.KmscError):
        ctx.codec_decode(data[:-3])
    with pytest.raises(kmsc.KmscError):
        ctx.codec_decode(b"XXXX" + data[4:])
    # bucket sizes that do not add up to n_keys
    bad = bytearray(data)
    bad[48 + (1 << 14) // 4] ^= 0x7F
    with pytest.raises(kmsc.KmscError):
        ctx.codec_decode(bytes(bad))
