"""N>1 host logic on CPU: world_size-2 gloo run of the prefix-sharded matrix sum.
Each rank takes its bucket range (shard.bucket_cuts), computes the partial N x N
intersection matrix over that range (here with the oracle, the GPU path is covered
by -m gpu tests) and one all-reduce must reproduce the full matrix."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
sys.path.insert(0, str(ROOT / "tests"))


def _worker(rank, world, port, out_dir):
    import shard
    import synth
    from _oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    K, N, kb = 23, 14, 4
    seqs = synth.phylogeny_sequences(6, 30000, p=0.01, seed=5)
    sets = [synth.kmer_set_of(s, K) for s in seqs]
    csr = [synth.csr_of(km, K, N, kb) for km in sets]
    cuts = shard.bucket_cuts(csr[0][0], world)
    lo, hi = cuts[rank], cuts[rank + 1]
    ids = np.arange(lo, hi, dtype=np.int32)
    part, visits = o.pair_counts([c[0] for c in csr], [c[1] for c in csr], kb, 1 << N, bucket_ids=ids)
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    v = torch.tensor([visits], dtype=torch.int64)
    dist.all_reduce(v)
    if rank == 0:
        full, full_visits = o.pair_counts([c[0] for c in csr], [c[1] for c in csr], kb, 1 << N)
        ok = bool(np.array_equal(t.numpy(), full)) and int(v[0]) == full_visits
        # balance: every rank within 25 % of the mean key count
        sizes = [int(csr[0][0][cuts[r + 1]] - csr[0][0][cuts[r]]) for r in range(world)]
        ok = ok and max(sizes) <= 1.25 * (sum(sizes) / world)
        Path(out_dir, "result.txt").write_text("ok" if ok else f"mismatch {sizes}")
    dist.destroy_process_group()


def test_sharded_sum_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "result.txt").read_text() == "ok"


def test_bucket_cuts_properties():
    import shard
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 100, 1 << 10)
    counts[:50] = 0
    offs = np.concatenate([[0], np.cumsum(counts)])
    for world in (1, 2, 3, 4, 8):
        c = shard.bucket_cuts(offs, world)
        assert c[0] == 0 and c[-1] == 1 << 10 and all(a <= b for a, b in zip(c, c[1:])) and len(c) == world + 1
    assert shard.bucket_cuts(np.zeros(17, np.int64), 4) == [0, 0, 0, 0, 16]
