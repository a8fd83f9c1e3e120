"""Generates tests/golden/ref_golden.json by running the REFERENCE'S OWN CODE
(oracle/_ref/libkmsc_ref.so = /root/reference/lib headers, unmodified, compiled by
oracle/Makefile) on seeded inputs. Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there read the committed JSON.
"""
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from _oracle import CONFIGS, Oracle, Ref, set_ref_seed  # noqa: E402

SEED = 20261018


def randseq(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


def mutate(rng, s, p):
    a = np.array(list(s))
    m = rng.random(len(a)) < p
    a[m] = rng.choice(list("ACGT"), int(m.sum()))
    return "".join(a)


def main():
    set_ref_seed(SEED)
    r, o = Ref(), Oracle()
    rng = np.random.default_rng(SEED)
    g = {"seed": SEED, "configs": {str(k): list(v) for k, v in CONFIGS.items()}}

    # --- Kmer ops (kmer.h) -------------------------------------------------
    g["kmer"] = []
    for cfg, (K, N, kb) in CONFIGS.items():
        for _ in range(8):
            s = randseq(rng, K)
            bits, _ = r.kmer_op(cfg, 0, s=s)
            comp, comp_s = r.kmer_op(cfg, 1, s=s)
            can, can_s = r.kmer_op(cfg, 2, s=s)
            c = str(rng.choice(list("ACGT")))
            nxt, _ = r.kmer_op(cfg, 3, s=s, c=c)
            prv, _ = r.kmer_op(cfg, 4, s=s, c=c)
            b, key, back = r.bucket_key(cfg, can)
            assert back == can
            g["kmer"].append(dict(cfg=cfg, s=s, bits=bits, comp=comp, comp_s=comp_s, can=can, can_s=can_s,
                                  c=c, next=nxt, prev=prv, bucket=b, key=key))

    # --- KmerCounter (kmer_counter.h) ----------------------------------------
    g["counter"] = []
    g["counter_reads"] = []
    for cfg in (0, 2, 4, 5):
        K = CONFIGS[cfg][0]
        base = randseq(rng, 400)
        reads = []
        for i in range(30):
            a = int(rng.integers(0, 300))
            rd = mutate(rng, base[a:a + int(rng.integers(K - 2, 100))], 0.02)
            if rng.random() < 0.3 and len(rd) > 3:
                p = int(rng.integers(0, len(rd)))
                rd = rd[:p] + "N" + rd[p + 1:]
            reads.append(rd)
        reads.append("N" * 5)
        reads.append("")
        g["counter_reads"].append(reads)
        for canonical in (False, True):
            for cutoff in (1, 2, 3):
                kmers, counts, kept, cut = r.count_reads(cfg, reads, canonical, cutoff)
                g["counter"].append(dict(cfg=cfg, reads_id=len(g["counter_reads"]) - 1, canonical=canonical, cutoff=cutoff,
                                         kmers=kmers.tolist(), counts=counts.tolist(),
                                         kept=kept.tolist(), cutoff_count=cut))
    # saturation: one k-mer 300 times
    K = 5
    reads = ["ACGTA"] * 300 + ["CCCCC"] * 255 + ["GGGGG"] * 254
    g["counter_reads"].append(reads)
    kmers, counts, kept, cut = r.count_reads(0, reads, False, 255)
    g["counter"].append(dict(cfg=0, reads_id=len(g["counter_reads"]) - 1, canonical=False, cutoff=255, kmers=kmers.tolist(),
                             counts=counts.tolist(), kept=kept.tolist(), cutoff_count=cut))
    g["fasta"] = []
    for lines in ([">a", "ACGTN", ">b", "GGGTTACA"], [">a", "ACGT", ">b"], ["a", "ACGT"], [">a", "ACGu"],
                  [">a", "acgt"], ["", "ACGT"], [], [">x", ""]):
        rc, nd = r.fasta(0, lines, True)
        g["fasta"].append(dict(lines=lines, rc=rc, n_distinct=nd if rc == 0 else None))

    # --- KmerSetCompact (kmer_set_compact.h) + GetKmerSetFromSPSS ----------------
    g["compact"] = []
    g["compact_strings"] = []
    for cfg in (1, 2, 3, 4, 5):
        K, N, kb = CONFIGS[cfg]
        strs = [randseq(rng, int(rng.integers(K, 160))) for _ in range(10)]
        strs.append(strs[0])  # duplicate k-mers on purpose
        g["compact_strings"].append(strs)
        nb = 1 << N
        for ids in (rng.permutation(nb)[: max(1, nb // 50)].astype(np.int32),
                    np.arange(nb, dtype=np.int32)[::-1].copy()):
            for canonical in (True, False):
                offs, keys, size, weight = r.sampled_set(cfg, strs, canonical, ids)
                sset, h = r.set_from_spss(cfg, strs, canonical)
                cnt = np.diff(offs)
                nz = np.nonzero(cnt)[0]
                g["compact"].append(dict(cfg=cfg, strings_id=len(g["compact_strings"]) - 1, canonical=canonical,
                                         bucket_ids=("reversed_all" if len(ids) == nb else ids.tolist()),
                                         counts_nz=[[int(i), int(cnt[i])] for i in nz],
                                         keys=keys.tolist(), size=size, weight=weight,
                                         set_size=len(sset), set_hash=h,
                                         set_head=sset[:16].tolist()))

    # --- KmerSet algebra (kmer_set.h) ----------------------------------------------
    g["setops"] = []
    for cfg in (0, 2, 4, 5):
        K = CONFIGS[cfg][0]
        hi = min(1 << (2 * K), 1 << 62)
        pool = np.unique(rng.integers(0, hi, 200, dtype=np.uint64)) if 2 * K > 12 else np.arange(1 << (2 * K), dtype=np.uint64)
        a = np.sort(rng.choice(pool, len(pool) // 2, replace=False))
        b = np.sort(rng.choice(pool, len(pool) // 2, replace=False))
        g["setops"].append(dict(cfg=cfg, a=a.tolist(), b=b.tolist(),
                                add=r.set_op(cfg, "add", a, b).tolist(),
                                sub=r.set_op(cfg, "sub", a, b).tolist(),
                                intersection=r.set_op(cfg, "intersection", a, b).tolist(),
                                diff=r.set_op(cfg, "diff", a, b), hash=r.set_op(cfg, "hash", a, b)))

    # --- KmerSetSet (kmer_set_set.h) seeded, 1 worker ----------------------------------
    g["kmersetset"] = []
    for cfg, glen, nsets in ((2, 3000, 6), (4, 2500, 5), (3, 2000, 9)):
        K, N, kb = CONFIGS[cfg]
        genome = randseq(rng, glen)
        seqs = [genome]
        for i in range(1, nsets):
            seqs.append(mutate(rng, seqs[(i - 1) // 2], 0.01))
        with tempfile.TemporaryDirectory() as d:
            files, spss_all, kmer_sets = [], [], []
            for i, s in enumerate(seqs):
                kmers = o.set_from_spss([s], K, True)
                spss, w = r.spss_from_set(cfg, kmers, True)
                p = os.path.join(d, f"in{i}.txt")
                Path(p).write_text("".join(x + "\n" for x in spss))
                files.append(p)
                spss_all.append(spss)
                kmer_sets.append(kmers)
            c0 = r.seed_counter()
            out = r.kmer_set_set(cfg, files, True, n_workers=1, dump_dir=os.path.join(d, "dump"))
            assert out["rc"] == 0
            ids = r.random_ints(c0, (1 << N) // 50, 0, (1 << N) - 1)
            merges = []
            for line in out["log"].split("\n"):
                if line.startswith("j = "):
                    parts = line.replace(",", "").split()
                    merges.append([int(parts[2]), int(parts[5]), int(parts[8])])
            rd = [r.reader_get(cfg, os.path.join(d, "dump"), True, i)[:2] for i in range(nsets)]
            meta = Path(d, "dump", "meta.txt").read_text().split("\n")
            g["kmersetset"].append(dict(cfg=cfg, spss=spss_all, bucket_ids=ids.tolist(), merges=merges,
                                        n_nodes=out["n_nodes"], sizes=out["sizes"].tolist(),
                                        hashes=[int(h) for h in out["hashes"]],
                                        reader=[[int(a), int(b)] for a, b in rd],
                                        in_sizes=[len(k) for k in kmer_sets],
                                        in_hashes=[int(o.set_hash(k)) for k in kmer_sets], meta=meta[:2]))

    # --- ParallelDisjointSet -----------------------------------------------------------
    g["dsu"] = []
    for n in (10, 200):
        d = r.lib.ref_dsu_new(n)
        pairs = rng.integers(0, n, (n, 2)).tolist()
        for x, y in pairs:
            r.lib.ref_dsu_unite(d, x, y)
        roots = [r.lib.ref_dsu_find(d, i) for i in range(n)]
        r.lib.ref_dsu_free(d)
        g["dsu"].append(dict(n=n, pairs=pairs, roots=roots))

    out = HERE / "ref_golden.json"
    out.write_text(json.dumps(g, separators=(",", ":")))
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
