"""Multi-GPU inside the library (kmsc_comm_*, kmsc_sets_exchange, all-reduce inside kmsc_pair_counts):
 * one rank: the communicator API and the exchange with itself (runs on the driver's single GPU);
 * two ranks over NCCL when the box has two GPUs (gpurun --gpus 2): tests/multi_rank_check.py."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))

pytestmark = pytest.mark.gpu


def test_single_rank_communicator_and_exchange(oracle):
    import kmsc
    import synth
    from test_gpu_decode_batch import pack
    ctx = kmsc.Context(0)
    try:
        ctx.comm_init(0, 1, kmsc.Context.comm_unique_id())
    except kmsc.KmscError as e:
        if "NCCL is not available" in str(e):
            pytest.skip("libnccl.so.2 not loadable here")
        raise
    assert ctx.comm_info() == (0, 1)
    K, N, KB = 23, 14, 4
    seqs = synth.phylogeny_sequences(3, 20000, p=0.01, seed=9)
    sets = [synth.kmer_set_of(s, K, True) for s in seqs]
    packed = [pack([synth.to_ascii(s).decode()]) for s in seqs]
    mine = ctx.sets_from_packed_batch(K, N, KB, [p[0] for p in packed], [p[1] for p in packed])
    got = ctx.sets_exchange(mine, [0, 1 << N], 3)
    for s, want in zip(got, sets):
        assert np.array_equal(s.to_kmers(), want)
    W = ctx.pair_counts(got)
    csr = [oracle.to_csr(s, K, N, KB) for s in sets]
    want_w, _ = oracle.pair_counts([c[0] for c in csr], [c[1] for c in csr], KB, 1 << N)
    iu = np.triu_indices(3, 1)
    assert np.array_equal(W[iu], want_w[iu])
    with pytest.raises(kmsc.KmscError):
        ctx.sets_exchange(mine, [0, 100], 3)          # cuts must cover [0, 2^N]
    with pytest.raises(kmsc.KmscError):
        ctx.comm_init(0, 1, kmsc.Context.comm_unique_id())   # already has one
    ctx.comm_destroy()
    with pytest.raises(kmsc.KmscError):
        ctx.sets_exchange(mine, [0, 1 << N], 3)       # no communicator
    ctx.close()


def test_two_ranks_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(ROOT / "tests" / "multi_rank_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI-RANK OK 2" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
