"""Host side (C++17 facade + drop-in executables) on a GPU:
 * tests/cpp/facade_test.cc restates the reference's own gtest cases (test/kmer.cc,
   kmer_set.cc, kmer_counter.cc, kmer_set_compact.cc, kmer_set_set.cc,
   parallel_disjoint_set.cc, spss.cc) against the kmsc classes of the same names;
 * the three executables run end to end: kmerset-build -> kmerset-multiple-compress ->
   kmerset-multiple-decompress, and the printed (size, XOR hash) of every reconstructed
   set must equal the oracle's for the original input (the reference's end-to-end
   observable, src/kmerset-multiple-decompress.cc:57-79)."""
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "kmer-sets-compression_b200" / "host"
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bins():
    r = subprocess.run(["make", "-s", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return HOST / "bin"


def test_facade_cases(bins):
    r = subprocess.run([str(bins / "facade_test")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("K", [15, 23])
def test_cli_end_to_end(bins, oracle, tmp_path, K):
    import synth
    rng = np.random.default_rng(K)
    seqs = synth.phylogeny_sequences(5, 6000, p=0.02, seed=K)
    spss_files, want = [], []
    for i, s in enumerate(seqs):
        # reads covering the sequence 3x -> FASTA -> kmerset-build (cutoff 2 drops nothing that is covered twice)
        txt = synth.to_ascii(s).decode()
        fasta = tmp_path / f"in{i}.fa"
        lines = []
        for rep in range(3):
            for a in range(0, len(txt) - 200, 150):
                lines += [f">r{rep}_{a}", txt[a:a + 230]]
        fasta.write_text("\n".join(lines) + "\n")
        reads = lines[1::2]
        kmers, counts = oracle.count_reads(reads, K, True)
        kept, _ = oracle.counter_to_set(kmers, counts, 2)
        want.append((len(kept), oracle.set_hash(kept)))
        out = tmp_path / f"set{i}.txt"
        # K = 23 streams the file in ~20 KB chunks of whole records (the config-4 path)
        extra = ["--chunk_bytes=20000"] if K == 23 else []
        r = subprocess.run([str(bins / "kmerset-build"), f"--k={K}", "--cutoff=2", "--check", f"--out={out}", str(fasta)] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert f"kmer_set.Size() = {len(kept)}" in r.stderr and f"kmer_set.Hash() = {oracle.set_hash(kept)}" in r.stderr
        assert "kmer_set_compact -> KmerSet: ok" in r.stderr
        spss_files.append(str(out))
    outdir = tmp_path / "dump"
    r = subprocess.run([str(bins / "kmerset-multiple-compress"), f"--k={K}", f"--out={outdir}", "--seed=3",
                        f"--out_graph={tmp_path / 'g.dot'}"] + spss_files, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    meta = (outdir / "meta.txt").read_text().split("\n")
    n_nodes = int(meta[1])
    assert n_nodes >= 5 and (outdir / f"{n_nodes - 1}.txt").exists()
    assert (tmp_path / "g.dot").read_text().startswith("digraph G {")
    r = subprocess.run([str(bins / "kmerset-multiple-decompress"), f"--k={K}", "--n=5", str(outdir)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    hashes = [int(x) for x in re.findall(r"kmer_set.Hash\(\) = (\d+)", r.stderr)]
    sizes = [int(x) for x in re.findall(r"kmer_set.Size\(\) = (\d+)", r.stderr)]
    assert list(zip(sizes, hashes)) == want


def test_build_reader_modes(bins, oracle, tmp_path):
    """f3: the file side of kmerset-build -- page-locked reader thread (default), pageable overlapped reader,
    read-and-count in turn, and a piped decompressor -- gives the same set on chunks of a few records; a file
    with an odd number of lines is refused with the reference's message by every mode"""
    import os
    import synth
    K = 23
    s = synth.random_genome(9000, 5)
    txt = synth.to_ascii(s).decode()
    lines = []
    for rep in range(3):
        for a in range(0, len(txt) - 200, 97):
            lines += [f">r{rep}_{a}", txt[a:a + 211]]
    fasta = tmp_path / "reads.fa"
    fasta.write_text("\n".join(lines) + "\n")
    kmers, counts = oracle.count_reads(lines[1::2], K, True)
    kept, _ = oracle.counter_to_set(kmers, counts, 2)
    want = (f"kmer_set.Size() = {len(kept)}", f"kmer_set.Hash() = {oracle.set_hash(kept)}")
    odd = tmp_path / "odd.fa"
    odd.write_text("\n".join(lines[:-1]) + "\n")
    for mode, extra in [(None, []), ("1", []), ("0", []), (None, ["--decompressor=cat"])]:
        env = dict(os.environ)
        env.pop("KMSC_IO_OVERLAP", None)
        if mode is not None:
            env["KMSC_IO_OVERLAP"] = mode
        for chunk in (3000, 10_000_000):
            r = subprocess.run([str(bins / "kmerset-build"), f"--k={K}", "--cutoff=2", f"--chunk_bytes={chunk}", str(fasta)] + extra,
                               capture_output=True, text=True, timeout=600, env=env)
            assert r.returncode == 0, r.stderr
            assert want[0] in r.stderr and want[1] in r.stderr, (mode, extra, chunk, r.stderr[-400:])
        r = subprocess.run([str(bins / "kmerset-build"), f"--k={K}", "--chunk_bytes=3000", str(odd)] + extra,
                           capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode != 0 and "even number of lines" in r.stderr, (mode, r.stderr[-300:])
