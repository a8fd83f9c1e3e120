#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json):
pairwise k-mer intersections per second, as key-visits/s, where

    key-visits = sum_{i<j} sum_{b in B} (len_i[b] + len_j[b])

is the work the reference's GetEdgeWeight merge performs for the same matrix
(lib/core/kmer_set_set.h:158-219).

Workload at N=1 = BASELINE configs[1] ("C2"): 64 sets x 10M canonical 23-mers
(<23,14,uint32>), sets derived from one random genome by a binary phylogeny of
0.2 % substitutions, exact all-bucket N x N matrix. With N>1 ranks the k-mer prefix
space is sharded by cumulative key count, every rank holds all 64 sets restricted
to its prefix range and the genome grows with N so the per-GPU key count stays
fixed ("weak" scaling); the partial matrices are summed by one NCCL all-reduce.

  value  : key-visits/s with the CSR sets resident in HBM (P3 plan + main kernel
           [+ all-reduce]), CUDA events on the launching stream, max over ranks.
  e2e    : same metric through the C ABI from HOST buffers: per step the 2-bit
           packed SPSS of every set (what KmerSetCompact holds in memory) is
           copied from pinned host memory, decoded to CSR on the device (P2), the
           matrix computed (P3) and read back.
  --impl reference : the reference's own unmodified headers (oracle/_ref) running
           KmerSetSet's constructor up to "calculated initial weights"
           (GetSampledKmerSet per set + the all-pairs GetEdgeWeight loop over its
           own 2 % bucket sample) on a bounded number of sets, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
sys.path.insert(0, str(ROOT / "tests"))

K, N, KB = 23, 14, 4
METRIC = "pairwise k-mer intersection throughput (key-visits/s = sum over pairs and buckets of len_i + len_j)"
UNIT = "key-visits/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sets", type=int, default=64)
    ap.add_argument("--kmers", type=int, default=10_000_000, help="k-mers per set per GPU")
    ap.add_argument("--p", type=float, default=0.002)
    ap.add_argument("--ref-sets", type=int, default=0, help="sets in the reference sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stage", action="store_true")
    ap.add_argument("--no-allreduce", action="store_true", help="diagnostic: time the per-rank partial matrices only")
    ap.add_argument("--emulate", default="", help="diagnostic, one GPU: R/W = the shard rank R of a W-rank job would hold")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# synthetic data (torch on the GPU for speed; numpy fallback for the CPU arm)
# ----------------------------------------------------------------------------

def gen_sequences_torch(n_sets, G, p, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(12345)
    seqs = [torch.randint(0, 4, (G,), dtype=torch.uint8, device=device, generator=g)]
    for i in range(1, n_sets):
        g.manual_seed(1000 + i)
        par = seqs[(i - 1) // 2]
        m = torch.rand(G, device=device, generator=g) < p
        d = torch.randint(1, 4, (G,), dtype=torch.uint8, device=device, generator=g)
        seqs.append(torch.where(m, (par + d) & 3, par))
    return seqs


def pack_torch(codes):
    """codes uint8 (0..3) -> int64 words, 32 bases per word, first base in the top bits"""
    import torch
    G = codes.numel()
    nw = (G + 31) // 32
    pad = torch.zeros(nw * 32, dtype=torch.int64, device=codes.device)
    pad[:G] = codes.to(torch.int64)
    shifts = (62 - 2 * torch.arange(32, device=codes.device, dtype=torch.int64))
    words = (pad.view(nw, 32) << shifts).sum(dim=1)
    return torch.cat([words, torch.zeros(2, dtype=torch.int64, device=codes.device)])


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every
    few milliseconds (the timed region of a short run lasts tens of milliseconds)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mx, self.reasons = [], None, set()
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            dev = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(dev.split(",")[gpu_index]) if dev and dev.split(",")[gpu_index].strip().isdigit() else gpu_index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _poll(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ----------------------------------------------------------------------------
# reference arm / cpu baseline
# ----------------------------------------------------------------------------

def reference_run(n_sets_sample, G, p, steps, warmup, n_workers):
    """Times the reference's own constructor phases on `n_sets_sample` sets of the workload."""
    import synth
    from _oracle import Ref, set_ref_seed
    set_ref_seed(4242)
    ref = Ref()
    ref.lib.ref_set_log_level(4)
    seqs = synth.phylogeny_sequences(n_sets_sample, G, p)
    tmp = tempfile.mkdtemp(prefix="kmsc_ref_")
    files, lens = [], []
    for i, s in enumerate(seqs):
        f = os.path.join(tmp, f"{i}.txt")
        with open(f, "wb") as fh:
            for piece in synth.split_strings(s, K, 100000):  # an SPSS-like multi-string file
                fh.write(piece + b"\n")
        files.append(f)
        km = synth.kmers_of(s, K, True)  # duplicates kept, like GetSampledKmerSet
        lens.append(np.bincount((km >> np.uint64(2 * K - N)).astype(np.int64), minlength=1 << N))
    lens = np.stack(lens)
    times, visits = [], []
    for it in range(warmup + steps):
        c0 = ref.seed_counter()
        out = ref.kmer_set_set(4, files, True, n_workers=n_workers, stop_after_weights=True)
        assert out["rc"] == 1, out["rc"]
        ids = ref.random_ints(c0, (1 << N) // 50, 0, (1 << N) - 1)
        v = (n_sets_sample - 1) * int(lens[:, ids].sum())
        if it >= warmup:
            times.append(out["phase_s"][0] + out["phase_s"][1])
            visits.append(v)
    for f in files:
        os.remove(f)
    os.rmdir(tmp)
    return sum(visits) / sum(times), float(np.mean(times)), times


def port_run(n_sets_sample, G, p, n_threads):
    """fallback CPU baseline: the oracle's merge loop (kmsc_oracle.c) on all buckets"""
    import synth
    from _oracle import Oracle
    o = Oracle()
    seqs = synth.phylogeny_sequences(n_sets_sample, G, p)
    offs_l, keys_l = [], []
    for s in seqs:
        offs, keys = synth.csr_of(synth.kmer_set_of(s, K), K, N, KB)
        offs_l.append(offs)
        keys_l.append(keys)
    t = time.time()
    _, v = o.pair_counts(offs_l, keys_l, KB, 1 << N, n_threads=n_threads)
    dt = time.time() - t
    return v / dt, dt


def main():
    args = parse_args()
    # stdout carries exactly ONE line (the JSON): anything a library prints there (NCCL's version
    # banner, for one) goes to stderr instead
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    G = args.kmers + K - 1  # bases per set per GPU
    cores = os.cpu_count() or 1
    config = {"workload": f"C2: {args.sets} sets x {args.kmers} canonical {K}-mers per GPU (<{K},{N},uint32>), "
                          f"binary phylogeny p={args.p}, all-pairs intersection matrix over all {1 << N} buckets (exact)",
              "n_sets": args.sets, "kmers_per_set_per_gpu": args.kmers, "k": K, "bucket_bits": N,
              "parallelism": f"prefix-sharded x{world}" if world > 1 else "single GPU",
              "l2": "inputs (2.56 GB/GPU) exceed the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        from _oracle import Ref
        n_s = args.ref_sets or (8 if args.steps + args.warmup <= 16 else 4 if args.steps + args.warmup <= 40 else 2)
        n_s = min(n_s, args.sets)
        if Ref.available():
            val, mean_s, _ = reference_run(n_s, G, args.p, args.steps, args.warmup, cores)
            kind = "reference"
            sample = (f"first {n_s} of the {args.sets} sets; KmerSetSet constructor up to 'calculated initial weights' "
                      f"(GetSampledKmerSet + all-pairs GetEdgeWeight) over the reference's own 2% bucket sample "
                      f"({(1 << N) // 50} of {1 << N} buckets), n_workers={cores}")
        else:
            val, mean_s = port_run(n_s, G, args.p, cores)
            kind = "port"
            sample = f"first {n_s} sets, all buckets, oracle merge loop, {cores} threads"
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), file=out, flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import kmsc
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one explicit (non-default) stream shared by torch (events, NCCL ordering, copies) and
    # libkmsc, so CUDA events time the launching stream and collectives are ordered with it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = kmsc.Context(local_rank, stream.cuda_stream)

    # data: genome grows with the number of ranks; every rank builds the same sequences
    shard_rank, shard_world = rank, world
    if args.emulate and world == 1:
        shard_rank, shard_world = (int(x) for x in args.emulate.split("/"))
    Gtot = args.kmers * shard_world + K - 1
    seqs = gen_sequences_torch(args.sets, Gtot, args.p, dev)
    str_offs = np.array([0, Gtot], np.int64)
    pinned, nbytes_in = [], 0
    for s in seqs:
        w = pack_torch(s)
        h = torch.empty(w.numel(), dtype=torch.int64, pin_memory=True)
        h.copy_(w)
        pinned.append(h)
        nbytes_in += (Gtot + 31) // 32 * 8 + str_offs.nbytes
    del seqs
    torch.cuda.synchronize()

    # prefix shard by cumulative key count of set 0
    lo, hi = 0, 1 << N
    if shard_world > 1:
        s0 = ctx.set_from_packed(K, N, KB, None, str_offs, words_ptr=pinned[0].data_ptr())
        offs, _ = s0.to_csr()
        s0.free()
        cuts = [int(np.searchsorted(offs, offs[-1] * r / shard_world)) for r in range(shard_world + 1)]
        cuts[0], cuts[-1] = 0, 1 << N
        lo, hi = cuts[shard_rank], cuts[shard_rank + 1]

    def build_sets():
        # one batched launch sequence for all sets (the per-set loop of KmerSetSet's constructor)
        return ctx.sets_from_packed_batch(K, N, KB, None, [str_offs] * len(pinned), bucket_lo=lo, bucket_hi=hi,
                                          words_ptrs=[h.data_ptr() for h in pinned])

    sets = build_sets()
    n = len(sets)
    d_out = torch.zeros(n * n, dtype=torch.int64, device=dev)
    keys_local = sum(s.n_keys for s in sets)
    visits_local = (n - 1) * keys_local

    def step_resident():
        ctx.pair_counts_device(sets, d_out.data_ptr())
        if world > 1 and not args.no_allreduce:
            dist.all_reduce(d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main_ms, plan_ms, algo_bytes, main_launches = 0.0, 0.0, 0.0, 0
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        st = ctx.pair_counts_stats()
        main_ms += st["main_ms"]; plan_ms += st["plan_ms"]; algo_bytes += st["algo_bytes"]
        main_launches += int(st["main_launches"])
    e1.record(stream)
    barrier()
    p3_build = ctx.pair_counts_build()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, float(visits_local)], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # every rank's own step and kernel time (the line's ms_per_step is the max over ranks)
        mine = torch.tensor([ms / args.steps, main_ms / max(1, main_launches), float(keys_local)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(x[0]), 4) for x in allr], "kernel_ms": [round(float(x[1]), 4) for x in allr],
                    "keys": [int(x[2]) for x in allr]}
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, visits_total = float(tm[0]), float(ts[1])
    else:
        visits_total = float(visits_local)
    value = visits_total * args.steps / (ms / 1e3)
    W = d_out.cpu().numpy().reshape(n, n)

    # ---- "pairwise-weight plus diff" stage: the matrix + one split per spanning-tree edge ----
    # (north_star: MST over d(i,j) = |Si| + |Sj| - 2 W[i][j]; per edge the two difference sets
    # via kmsc_pair_split). Bytes are the algorithmic B_w + sum B_s of SURVEY 8(d).
    stage = None
    if not args.no_stage:
        sizes = np.diag(W).astype(np.int64)
        dist_m = sizes[:, None] + sizes[None, :] - 2 * W
        in_tree, parent, best = np.zeros(n, bool), np.zeros(n, np.int64), np.full(n, np.iinfo(np.int64).max)
        in_tree[0] = True
        best[1:], parent[1:] = dist_m[0, 1:], 0
        edges = []
        for _ in range(n - 1):  # Prim; ties -> smallest index
            cand = np.where(in_tree, np.iinfo(np.int64).max, best)
            c = int(np.argmin(cand))
            edges.append((int(parent[c]), c))
            in_tree[c] = True
            upd = (~in_tree) & (dist_m[c] < best)
            best[upd], parent[upd] = dist_m[c][upd], c
        barrier()
        js, ks = [sets[pa] for pa, _ in edges], [sets[ch] for _, ch in edges]
        W_loc = W
        if world > 1:  # W is the all-reduced matrix; the hints are this rank's partial counts
            d_loc = torch.zeros(n * n, dtype=torch.int64, device=dev)
            ctx.pair_counts_device(sets, d_loc.data_ptr())
            W_loc = d_loc.cpu().numpy().reshape(n, n)
        hint = np.array([W_loc[pa, ch] for pa, ch in edges], np.int64)
        for _ in range(2):  # warm-up (allocation pool, first-launch attributes)
            _, op, oc = ctx.pair_split_batch(js, ks, inter_hint=hint, want_inter=False)
            for s in op + oc:
                s.free()
        barrier()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        split_reps = 5
        es0.record(stream)
        for _ in range(split_reps):
            _, op, oc = ctx.pair_split_batch(js, ks, inter_hint=hint, want_inter=False)
            split_bytes = sum((a.n_keys + b.n_keys + x.n_keys + y.n_keys) * KB + 5 * ((1 << N) + 1) * 4
                              for a, b, x, y in zip(js, ks, op, oc))
            for s in op + oc:
                s.free()
        es1.record(stream)
        barrier()
        split_ms = es0.elapsed_time(es1) / split_reps
        w_ms = ms / args.steps
        w_bytes = algo_bytes / max(1, main_launches)
        stage = {"what": "all-pairs matrix + the two difference sets of each of the n-1 MST edges (one kmsc_pair_split_batch), per GPU",
                 "ms": w_ms + split_ms, "weights_ms": w_ms, "splits_ms": split_ms, "n_splits": len(edges),
                 "algorithmic_bytes": w_bytes + split_bytes,
                 "achieved_gbs": (w_bytes + split_bytes) / ((w_ms + split_ms) / 1e3) / 1e9}

    # ---- e2e: host packed SPSS -> device CSR -> matrix -> host ------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 5))
        host_out = np.zeros((n, n), np.int64)

        def step_e2e_single():
            ss = build_sets()
            ctx.pair_counts_device(ss, d_out.data_ptr())
            if world > 1:
                dist.all_reduce(d_out)
            host_out[:] = d_out.cpu().numpy().reshape(n, n)
            for s in ss:
                s.free()

        # N > 1: the decode is split over the ranks BY SET (rank r decodes sets r, r + N, ... over
        # all buckets), one all-to-all hands every rank the slice of every set that falls into its
        # prefix range (a bucket range of a sorted set is one contiguous key slice), and every
        # rank imports all n sets restricted to its range. Without the exchange every rank would
        # have to scan the whole text of every set.
        m_own = n // world if world > 1 and n % world == 0 else 0
        cuts_np = np.asarray(cuts, np.int32) if world > 1 else None

        def step_e2e_exchange():
            mine = [rank + j * world for j in range(m_own)]
            full = [ctx.set_from_packed(K, N, KB, None, str_offs, words_ptr=pinned[i].data_ptr()) for i in mine]
            ko = np.stack([ctx.set_bucket_offsets(s, cuts_np) for s in full])       # [m_own][world + 1]
            sizes = (ko[:, 1:] - ko[:, :-1]).astype(np.int64)                        # keys of own set j for rank q
            t_send_sz = torch.from_numpy(np.ascontiguousarray(sizes.T)).to(dev)      # [world][m_own], dest-major
            t_recv_sz = torch.empty_like(t_send_sz)
            dist.all_to_all_single(t_recv_sz, t_send_sz)
            recv_sz = t_recv_sz.cpu().numpy()                                        # [src][j]
            nbq = [int(cuts[q + 1] - cuts[q]) + 1 for q in range(world)]
            nb_me = nbq[rank]
            in_k = [int(sizes[:, q].sum()) for q in range(world)]
            out_k = [int(recv_sz[r].sum()) for r in range(world)]
            send_keys = torch.empty(max(1, sum(in_k)), dtype=torch.int32, device=dev)
            recv_keys = torch.empty(max(1, sum(out_k)), dtype=torch.int32, device=dev)
            send_offs = torch.empty(m_own * sum(nbq), dtype=torch.int32, device=dev)
            recv_offs = torch.empty(world * m_own * nb_me, dtype=torch.int32, device=dev)
            kp, op = 0, 0
            for q in range(world):
                for j, s in enumerate(full):
                    ctx.set_export_range(s, int(cuts[q]), int(cuts[q + 1]), int(ko[j, q]), int(ko[j, q + 1]),
                                         send_offs.data_ptr() + op * 4, send_keys.data_ptr() + kp * KB)
                    kp += int(sizes[j, q]); op += nbq[q]
            dist.all_to_all_single(recv_keys[:sum(out_k)], send_keys[:sum(in_k)], out_k, in_k)
            dist.all_to_all_single(recv_offs, send_offs, [m_own * nb_me] * world, [m_own * x for x in nbq])
            ss = [None] * n
            kp = 0
            for r in range(world):
                for j in range(m_own):
                    cnt = int(recv_sz[r, j])
                    ss[r + j * world] = ctx.set_import_range(K, N, KB, lo, hi, recv_offs.data_ptr() + (r * m_own + j) * nb_me * 4,
                                                             recv_keys.data_ptr() + kp * KB, cnt)
                    kp += cnt
            ctx.pair_counts_device(ss, d_out.data_ptr())
            dist.all_reduce(d_out)
            host_out[:] = d_out.cpu().numpy().reshape(n, n)
            for s in ss + full:
                s.free()

        step_e2e = step_e2e_exchange if m_own > 0 else step_e2e_single

        for s in sets:
            s.free()
        step_e2e()
        barrier()
        e0.record(stream)
        for _ in range(e2e_steps):
            step_e2e()
        e1.record(stream)
        barrier()
        ms_e = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms_e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt[0])
        assert np.array_equal(host_out, W), "e2e matrix differs from the resident run"
        e2e = {"value": visits_total * e2e_steps / (ms_e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(nbytes_in // world) if m_own > 0 else int(nbytes_in),
               "how": ("decode split by set over the ranks + one all-to-all of prefix slices (NCCL) + all-reduce of the partial matrices"
                       if m_own > 0 else "every rank decodes its prefix range of every set"),
               "d2h_bytes_per_step": int(n * n * 8), "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_launch_ms = main_ms / max(1, main_launches)
    achieved = (algo_bytes / max(1, main_launches)) / (per_launch_ms / 1e3) / 1e9 if per_launch_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch: from the committed ncu capture of this
    # workload (it cannot be measured outside a profiler); null for any other shape
    traffic = None
    kernel_name = "pair_counts_stream_kernel" if p3_build == 1 else "pair_counts_kernel"
    try:
        tr = json.loads((ROOT / "profiles" / "p3_traffic.json").read_text())["kernels"][kernel_name]
        if args.sets == 64 and args.kmers == 10_000_000 and K == 23:
            traffic = int(tr["dram_bytes_read"]) + int(tr["dram_bytes_write"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_name,
                "build": "warp-wide multiway merge (related sets)" if p3_build == 1 else "shared-memory hash table",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": algo_bytes / max(1, main_launches),
                "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": main_ms / ms if ms else None,
                "plan_ms_per_step": plan_ms / args.steps}

    if stage is not None:
        stage["frac_of_hbm_peak"] = stage["achieved_gbs"] / peak
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from _oracle import Ref
        try:
            if Ref.available():
                v, mean_s, _ = reference_run(4, G, args.p, 1, 0, cores)
                cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": "first 4 of the sets; reference KmerSetSet constructor up to 'calculated "
                                          f"initial weights' over its own 2% bucket sample, n_workers={cores}; "
                                          f"{mean_s:.1f} s"}
            else:
                v, dt = port_run(8, G, args.p, cores)
                cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first 8 sets, all buckets, oracle merge loop; {dt:.1f} s"}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {ex}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stage": stage, "cpu_baseline": cpu_baseline,
            "check": {"W01": int(W[0, 1]), "W_diag0": int(W[0, 0]), "keys_per_gpu": int(keys_local),
                      "p3_stats": ctx.pair_counts_stats() if args.no_e2e and args.no_stage else None, "buckets": [int(lo), int(hi)]}}
    if per_rank is not None:
        line["per_rank"] = per_rank
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
