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


def _container(K, N, kb, n_keys, sizes, key_vals, oracle, n_sdata=None, n_kdata=None):
    """a KMSC version-1 container put together by hand (layout: csrc/codec.cu)"""
    import struct
    nb = 1 << N

    def stream(vals):  # 2-bit codes (bytes - 1), control bytes first, little-endian data
        ctrl, data = bytearray((len(vals) + 3) // 4), bytearray()
        for i, v in enumerate(vals):
            c = 0 if v < (1 << 8) else 1 if v < (1 << 16) else 2 if v < (1 << 24) else 3
            ctrl[i >> 2] |= c << (2 * (i & 3))
            data += int(v).to_bytes(4, "little")[: c + 1]
        return bytes(ctrl) + bytes(data), len(data)

    s_enc, sd = stream(sizes)
    k_enc, kd = stream(key_vals)
    hdr = struct.pack("<IIIIIIQQQ", 0x43534D4B, 1, K, N, kb, 2 if kb == 8 else 1, n_keys,
                      sd if n_sdata is None else n_sdata, kd if n_kdata is None else n_kdata)
    return hdr + s_enc + k_enc


def test_codec_rejects_sizes_that_only_add_up_mod_2_32(ctx, oracle):
    """bucket sizes {0xFFFFFFFF, n_keys + 1, 0, ...} pass a 32-bit total check; a huge
    n_sdata / n_kdata pair wraps the 64-bit length sum. Both must be refused before any key is written."""
    import struct
    import kmsc
    K, N, kb, n_keys = 15, 14, 2, 5
    nb = 1 << N
    good = _container(K, N, kb, n_keys, [n_keys] + [0] * (nb - 1), [1, 2, 3, 4, 5], oracle)
    magic = struct.unpack("<I", good[:4])[0]
    ref = oracle.codec_encode(K, N, kb, np.array([0] + [n_keys] * nb, np.int64), np.array([1, 3, 6, 10, 15], np.uint16))
    assert struct.unpack("<I", ref[:4])[0] == magic and ref == good, "hand-made container differs from the oracle's encoder"
    assert ctx.codec_decode(good).Size() == n_keys
    sizes = [0xFFFFFFFF, n_keys + 1] + [0] * (nb - 2)
    assert sum(sizes) % (1 << 32) == n_keys
    with pytest.raises(kmsc.KmscError):
        ctx.codec_decode(_container(K, N, kb, n_keys, sizes, [1, 2, 3, 4, 5], oracle))
    # lengths that only match modulo 2^64
    true_sd, true_kd = nb, 5   # one data byte per small value
    big = (1 << 64) - 4096
    with pytest.raises(kmsc.KmscError):
        ctx.codec_decode(_container(K, N, kb, n_keys, [n_keys] + [0] * (nb - 1), [1, 2, 3, 4, 5], oracle,
                                    n_sdata=big, n_kdata=(true_sd + true_kd - big) % (1 << 64)))
    # too many keys for one block (the encoder's own limit)
    with pytest.raises(kmsc.KmscError):
        ctx.codec_decode(_container(K, N, 8, (1 << 29) + 5, [n_keys] + [0] * (nb - 1), [1, 2, 3, 4, 5], oracle))
