"""One rank of the multi-GPU parity check (launched by tests/test_gpu_multi.py through torchrun, or by
hand: torchrun --nproc-per-node 2 tests/multi_rank_check.py). Every rank decodes its share of the
sets, kmsc_sets_exchange re-shards them by k-mer prefix, kmsc_pair_counts all-reduces inside the
library; rank by rank the imported sets and the matrix are compared with the oracle."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    import kmsc
    import synth
    from _oracle import Oracle
    from test_gpu_decode_batch import pack
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = kmsc.Context(local)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(kmsc.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.comm_init(rank, world, idt.cpu().numpy().tobytes())
    assert ctx.comm_info() == (rank, world)
    o = Oracle()
    K, N, KB = 23, 14, 4
    n = 4 * world
    seqs = synth.phylogeny_sequences(n, 30000, p=0.01, seed=5)
    seqs[3] = seqs[3][:K + 5]                       # a tiny set
    full_sets = [synth.kmer_set_of(s, K, True) for s in seqs]
    # cuts by the cumulative key count of set 0, like the product's sharding
    offs0, _ = synth.csr_of(full_sets[0], K, N, KB)
    cuts = [int(np.searchsorted(offs0, offs0[-1] * r / world)) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, 1 << N
    mine_idx = [rank + j * world for j in range(n // world)]
    packed = [pack([synth.to_ascii(seqs[i]).decode()]) for i in mine_idx]
    mine = ctx.sets_from_packed_batch(K, N, KB, [p[0] for p in packed], [p[1] for p in packed])
    got = ctx.sets_exchange(mine, cuts, n)
    lo, hi = cuts[rank], cuts[rank + 1]
    for g, s in enumerate(got):
        b = full_sets[g] >> np.uint64(2 * K - N)
        want = full_sets[g][(b >= lo) & (b < hi)]
        assert np.array_equal(s.to_kmers(), want), f"rank {rank}: set {g} differs after the exchange"
    W = ctx.pair_counts(got)                        # all-reduced inside the library
    csr = [o.to_csr(s, K, N, KB) for s in full_sets]
    want_w, _ = o.pair_counts([c[0] for c in csr], [c[1] for c in csr], KB, 1 << N)
    iu = np.triu_indices(n, 1)
    assert np.array_equal(W[iu], want_w[iu]), f"rank {rank}: all-reduced matrix differs from the oracle"
    assert np.array_equal(np.diag(W), [len(s) for s in full_sets])
    rows = ctx.pair_counts_rows(got, [1, n - 1])
    assert np.array_equal(rows, W[[1, n - 1]])
    # the partial matrices add up to the full one
    d_loc = torch.zeros(n * n, dtype=torch.int64, device=dev)
    ctx.pair_counts_device(got, d_loc.data_ptr(), partial=True)
    ctx.sync()
    tot = d_loc.clone()
    dist.all_reduce(tot)
    assert np.array_equal(tot.cpu().numpy().reshape(n, n), W)
    ctx.comm_destroy()
    dist.barrier()
    if rank == 0:
        print("MULTI-RANK OK", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
