"""Prefix sharding of the k-mer space across ranks.

|S_i & S_j| = sum over buckets, exactly as the reference sums it
(lib/core/kmer_set_set.h:161-181), so each rank takes a contiguous bucket range
and the partial N x N matrices are added by one all-reduce. Ranges are cut by
cumulative key count, not by bucket count: canonical k-mers put about
44/31/19/6 % of the keys under first base A/C/G/T."""
from __future__ import annotations

import numpy as np


def bucket_cuts(offs: np.ndarray, world: int) -> list[int]:
    """offs: CSR bucket offsets (2^N + 1) of a representative set (or the sum over sets).
    Returns world + 1 bucket boundaries; rank r owns buckets [cuts[r], cuts[r+1])."""
    offs = np.asarray(offs, np.int64)
    nb = len(offs) - 1
    total = int(offs[-1])
    cuts = [0]
    for r in range(1, world):
        b = int(np.searchsorted(offs, total * r / world, side="left"))
        cuts.append(min(max(b, cuts[-1]), nb))
    cuts.append(nb)
    return cuts


def rank_range(offs: np.ndarray, world: int, rank: int) -> tuple[int, int]:
    c = bucket_cuts(offs, world)
    return c[rank], c[rank + 1]
