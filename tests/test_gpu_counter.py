"""P1 parity (GPU): kmsc_count_fasta / kmsc_count_reads / kmsc_count_get through the C ABI
vs the oracle's restatement of KmerCounter (reference lib/core/kmer_counter.h:64-264) and
the golden fixtures produced by the reference's own code."""
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


def test_reference_known_answer(ctx, oracle):  # test/kmer_counter.cc:64-91
    s, cut, nd = ctx.count_reads(5, 3, 2, b"AACCGTT\nAACCGTA\n", canonical=False, cutoff=1)
    assert nd == 4 and cut == 0
    for name, c in (("AACCG", 2), ("ACCGT", 2), ("CCGTT", 1), ("CCGTA", 1), ("AAAAA", 0)):
        assert ctx.count_get(oracle.bits(name)) == c
    s2, cut2, _ = ctx.count_reads(5, 3, 2, b"AACCGTT\nAACCGTA", canonical=False, cutoff=2)
    assert cut2 == 2 and [oracle.string(int(k), 5) for k in s2.to_kmers()] == ["AACCG", "ACCGT"]


def test_golden_counter(ctx):
    for e in GOLD["counter"]:
        K, N, _ = CONFIGS[e["cfg"]]
        reads = GOLD["counter_reads"][e["reads_id"]]
        data = "".join(r + "\n" for r in reads).encode()
        s, cut, nd = ctx.count_reads(K, N, KB[K], data, canonical=e["canonical"], cutoff=e["cutoff"])
        assert nd == len(e["kmers"]) and cut == e["cutoff_count"]
        assert s.to_kmers().tolist() == e["kept"]
        for k, c in list(zip(e["kmers"], e["counts"]))[:40]:
            assert ctx.count_get(k) == c


@pytest.mark.parametrize("K,N", [(15, 14), (23, 14), (31, 14), (19, 10)])
@pytest.mark.parametrize("canonical", [True, False])
def test_fasta_counts_vs_oracle(ctx, oracle, K, N, canonical):
    rng = np.random.default_rng(K + 7)
    base = "".join(rng.choice(list("ACGT"), 20000))
    reads, lines = [], []
    for i in range(3000):
        a = int(rng.integers(0, 19800))
        r = base[a:a + int(rng.integers(1, 200))]
        if rng.random() < 0.1:
            p = int(rng.integers(0, len(r)))
            r = r[:p] + "N" + r[p + 1:]
        reads.append(r)
        lines += [f">r{i}", r]
    reads.append("")
    lines += [">empty", ""]
    kmers, counts = oracle.count_reads(reads, K, canonical)
    for cutoff in (1, 2, 4):
        kept, cut = oracle.counter_to_set(kmers, counts, cutoff)
        data = ("\n".join(lines) + "\n").encode()
        s, gcut, nd = ctx.count_fasta(K, N, KB[K], data, canonical=canonical, cutoff=cutoff)
        assert nd == len(kmers) and gcut == cut
        assert np.array_equal(s.to_kmers(), kept)
    idx = rng.integers(0, len(kmers), 30)
    for i in idx:
        assert ctx.count_get(int(kmers[i])) == int(counts[i])
    # no trailing newline: same result (std::getline semantics, core/io.h:31-34)
    s2, _, nd2 = ctx.count_fasta(K, N, KB[K], "\n".join(lines[:-2]).encode(), canonical=canonical, cutoff=1)
    assert nd2 == len(kmers)


@pytest.mark.parametrize("K,N,glen,n_reads", [(9, 10, 30000, 10000), (15, 14, 3000, 3000), (19, 10, 3000, 3000), (31, 14, 2000, 6000)])
@pytest.mark.parametrize("canonical", [True, False])
def test_high_coverage_long_runs(ctx, oracle, K, N, glen, n_reads, canonical):
    """sequencing-read shape (C4): every finest bucket holds 60-4000 k-mer instances -- the copies of a few
    distinct k-mers. Short k: several distinct k-mers per run (shared-memory counting sort + rank inside the
    sub-bin); long k over a small genome: a run is the copies of ONE k-mer (sub-bin longer than the rank limit:
    bitonic network). Counts, saturation and the cutoff against the oracle."""
    rng = np.random.default_rng(K * 31 + N)
    base = "".join(rng.choice(list("ACGT"), glen))
    reads = []
    for i in range(n_reads):
        a = int(rng.integers(0, glen - 100))
        reads.append(base[a:a + 100])
    kmers, counts = oracle.count_reads(reads, K, canonical)
    assert counts.max() >= 60
    data = ("\n".join(reads) + "\n").encode()
    for cutoff in (1, 40):
        kept, cut = oracle.counter_to_set(kmers, counts, cutoff)
        s, gcut, nd = ctx.count_reads(K, N, KB[K], data, canonical=canonical, cutoff=cutoff)
        assert nd == len(kmers) and gcut == cut
        assert np.array_equal(s.to_kmers(), kept)
    for i in rng.integers(0, len(kmers), 40):
        assert ctx.count_get(int(kmers[i])) == int(counts[i])


def test_saturation_at_255(ctx, oracle):  # kmer_counter.h:28-38, test/kmer_counter.cc:12-16
    data = ("ACGTA\n" * 300 + "CCCCC\n" * 255 + "GGGGG\n" * 254).encode()
    s, cut, nd = ctx.count_reads(5, 3, 2, data, canonical=False, cutoff=255)
    assert nd == 3 and cut == 1
    assert ctx.count_get(oracle.bits("ACGTA")) == 255 and ctx.count_get(oracle.bits("GGGGG")) == 254
    assert sorted(oracle.string(int(k), 5) for k in s.to_kmers()) == ["ACGTA", "CCCCC"]
    # cutoff is a uint8 (kmer_counter.h:214): 256 wraps to 0 and keeps everything
    s, cut, _ = ctx.count_reads(5, 3, 2, data, canonical=False, cutoff=256)
    assert cut == 0 and s.Size() == 3


def test_fasta_validation(ctx):
    import kmsc
    for e in GOLD["fasta"]:
        data = "".join(l + "\n" for l in e["lines"]).encode()
        if e["rc"] == 0:
            s, cut, nd = ctx.count_fasta(5, 3, 2, data, canonical=True, cutoff=1)
            assert nd == e["n_distinct"]
        else:
            with pytest.raises(kmsc.KmscError) as ei:
                ctx.count_fasta(5, 3, 2, data, canonical=True, cutoff=1)
            want = "even number of lines" if e["rc"] == 1 else "invalid FASTA file"
            assert want in str(ei.value), (e["lines"], str(ei.value))


@pytest.mark.parametrize("K,N", [(15, 14), (31, 14), (5, 3)])
def test_streaming_counter_equals_one_shot(ctx, oracle, K, N):
    """config 4 shape: the file is fed in chunks of whole records; counts are merged on the device
    with saturating uint8 adds (kmer_counter.h:28-38, 105-126) and equal the one-shot result"""
    rng = np.random.default_rng(K)
    base = "".join(rng.choice(list("ACGT"), 3000))
    reads, lines = [], []
    for i in range(4000):
        a = int(rng.integers(0, 2800))
        r = base[a:a + int(rng.integers(1, 150))]
        if rng.random() < 0.05:
            p = int(rng.integers(0, len(r)))
            r = r[:p] + "N" + r[p + 1:]
        reads.append(r)
        lines += [f">r{i}", r]
    kmers, counts = oracle.count_reads(reads, K, True)
    assert counts.max() == 255 or K != 5        # K = 5 saturates: exercises the 255 cap across chunks
    cuts = [0, 1000, 1002, 3000, 7000, 8000]    # record-aligned (even) line offsets, uneven chunk sizes
    chunks = [("\n".join(lines[a:b]) + "\n").encode() for a, b in zip(cuts, cuts[1:]) if b > a]
    for cutoff in (1, 3):
        kept, cut = oracle.counter_to_set(kmers, counts, cutoff)
        s, gcut, nd = ctx.count_chunks(K, N, KB[K], chunks, canonical=True, cutoff=cutoff)
        assert nd == len(kmers) and gcut == cut
        assert np.array_equal(s.to_kmers(), kept)
        one, ocut, ond = ctx.count_fasta(K, N, KB[K], b"".join(chunks), canonical=True, cutoff=cutoff)
        assert (ocut, ond) == (gcut, nd) and np.array_equal(one.to_kmers(), s.to_kmers())
    # the merged counter answers point queries
    s, _, _ = ctx.count_chunks(K, N, KB[K], chunks, canonical=True, cutoff=1)
    got = ctx.count_last_counts(s.n_keys)
    assert np.array_equal(got, counts)
    for i in rng.integers(0, len(kmers), 20):
        assert ctx.count_get(int(kmers[i])) == int(counts[i])


def test_streaming_counter_errors_and_empty(ctx):
    import kmsc
    s, cut, nd = ctx.count_chunks(15, 14, 2, [], canonical=True, cutoff=1)
    assert s.Size() == 0 and nd == 0 and cut == 0
    with pytest.raises(kmsc.KmscError) as ei:   # a chunk that ends inside a record
        ctx.count_chunks(15, 14, 2, [b">a\nACGTACGTACGTACGTAAAA\n>b\n"], canonical=True, cutoff=1)
    assert "even number of lines" in str(ei.value)
    with pytest.raises(kmsc.KmscError) as ei:
        ctx.count_chunks(15, 14, 2, [b">a\nACGT\n", b">b\nACGu\n"], canonical=True, cutoff=1)
    assert "invalid FASTA file" in str(ei.value)


def test_streaming_counter_prefetch(ctx, oracle):
    """kmsc_counter_prefetch (the next chunk's copy overlaps the counting of the current one): same counts with
    and without it, with announcements that are not followed up (another pointer is added) and repeated ones"""
    import ctypes as C
    import kmsc
    K, N = 23, 14
    rng = np.random.default_rng(3)
    base = "".join(rng.choice(list("ACGT"), 30000))
    chunks, reads = [], []
    for c in range(5):
        lines = []
        for i in range(400):
            a = int(rng.integers(0, 29700))
            r = base[a:a + int(rng.integers(30, 250))]
            reads.append(r)
            lines += [f">c{c}r{i}", r]
        chunks.append(np.frombuffer(("\n".join(lines) + "\n").encode(), np.uint8).copy())
    kmers, counts = oracle.count_reads(reads, K, True)
    kept, cut = oracle.counter_to_set(kmers, counts, 2)
    for prefetch in (True, False):
        s, gcut, nd = ctx.count_chunks(K, N, 4, chunks, canonical=True, cutoff=2, fasta=True, prefetch=prefetch)
        assert nd == len(kmers) and gcut == cut and np.array_equal(s.to_kmers(), kept)
        s.free()
    # announcements out of order / never used
    L = kmsc.lib()
    c = C.c_void_p()
    assert L.kmsc_counter_create(ctx.h, K, N, 4, 1, C.byref(c)) == 0
    ptr = lambda a: C.c_void_p(a.ctypes.data)
    assert L.kmsc_counter_prefetch(ctx.h, c, ptr(chunks[3]), len(chunks[3])) == 0
    assert L.kmsc_counter_prefetch(ctx.h, c, ptr(chunks[1]), len(chunks[1])) == 0
    for i in (0, 1, 2):
        assert L.kmsc_counter_add_fasta(ctx.h, c, ptr(chunks[i]), len(chunks[i])) == 0
    assert L.kmsc_counter_prefetch(ctx.h, c, ptr(chunks[4]), len(chunks[4])) == 0
    assert L.kmsc_counter_prefetch(ctx.h, c, ptr(chunks[4]), len(chunks[4])) == 0
    for i in (3, 4):
        assert L.kmsc_counter_add_fasta(ctx.h, c, ptr(chunks[i]), len(chunks[i])) == 0
    h, cc, nd = C.c_void_p(), C.c_int64(), C.c_int64()
    assert L.kmsc_counter_finish(ctx.h, c, 2, C.byref(h), C.byref(cc), C.byref(nd)) == 0
    L.kmsc_counter_free(ctx.h, c)
    s = kmsc.DeviceSet(ctx, h.value)
    assert nd.value == len(kmers) and cc.value == cut and np.array_equal(s.to_kmers(), kept)
