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
  --impl reference : the reference's own unmodified headers (oracle/_ref) on the SAME sets,
           all host threads. Two like-for-like measurements:
           exact   -- what the GPU arm computes: GetSampledKmerSet over ALL buckets for
                      every set (one task per set, as kmer_set_set.h:138-153) + the
                      all-pairs two-pointer merge (:158-219). The decode is timed on one
                      wave of `cores` sets, each step times the pair loop on those sets,
                      and `value` is the key-visits/s of the whole 64-set job at the
                      measured phase rates (ceil(64 / cores) decode waves + all pairs).
           sampled -- the reference's own constructor up to "calculated initial weights"
                      (its 2 % bucket sample) on all 64 sets, once; the GPU arm reports
                      the same job (`sampled`) with a 2 % bucket list of its own.
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
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4", "C5"],
                    help="C2 (BASELINE configs[1], the metric's config): 64 sets x 10M 23-mers per GPU, weak scaling; "
                         "C3 (configs[2]): 256 sets x 10M 23-mers in all, prefix-sharded over the GPUs, strong scaling; "
                         "C4 (configs[3]): k-mer counting of synthetic FASTA reads, k=31, cutoff 4 (--bytes, default 2 GB slice); "
                         "C5 (configs[4]): k=15 dense-bitmap AND-popcount Gram of 512 sets (one GPU)")
    ap.add_argument("--bytes", type=int, default=2_000_000_000, help="C4: bytes of FASTA to count")
    ap.add_argument("--chunk-bytes", type=int, default=256_000_000, help="C4: bytes per streamed chunk")
    ap.add_argument("--sets", type=int, default=0, help="number of sets (default: 64 for C2, 256 for C3)")
    ap.add_argument("--kmers", type=int, default=10_000_000, help="k-mers per set (C2: per GPU)")
    ap.add_argument("--p", type=float, default=0.002)
    ap.add_argument("--no-ref-sampled", action="store_true", help="reference arm: skip the sampled-constructor run on all sets")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stage", action="store_true")
    ap.add_argument("--profile-e2e", action="store_true", help="diagnostic: per-phase times of the e2e step on stderr (synchronises between phases)")
    ap.add_argument("--e2e-pieces", type=int, default=4, help="multi-GPU e2e: pieces of a rank's sets whose exchange overlaps the decode of the next piece")
    ap.add_argument("--no-e2e-overlap", action="store_true", help="multi-GPU e2e: exchange after the whole decode instead of overlapping the two halves")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle check of W at full size")
    ap.add_argument("--no-c1", action="store_true", help="skip the bounded C1 end-to-end run of the executable")
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

def gen_codes(n_sets, G, p):
    """the workload's sequences as numpy code arrays (0..3). Generated with torch on the GPU when
    there is one (then they are the very sets the GPU arm uses), else with the numpy generator of
    synth.py (same distribution, other sets)."""
    try:
        import torch
        if torch.cuda.is_available():
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            return [s.cpu().numpy() for s in gen_sequences_torch(n_sets, G, p, dev)], "torch-cuda generator (the GPU arm's sets)"
    except Exception:
        pass
    import synth
    return synth.phylogeny_sequences(n_sets, G, p), "numpy generator (same distribution, other sets than the GPU arm)"


def write_spss(codes, path, piece=100000, k=K):
    """one sequence as an SPSS-like text file: overlapping pieces spelling the same k-mers"""
    import synth
    with open(path, "wb") as fh:
        for s in synth.split_strings(codes, k, piece):
            fh.write(s + b"\n")


C1_MERGES = 3


def c1_files(tmp):
    """BASELINE configs[0] (C1): 8 sets of ~1 M canonical 15-mers, 1 Mbp windows of one 2 Mbp genome, p = 0.005"""
    import synth
    files = []
    for i, s in enumerate(synth.window_sequences(8, 2_000_000, 1_000_000, 0.005)):
        f = os.path.join(tmp, f"c1_{i}.txt")
        write_spss(s, f, k=15)
        files.append(f)
    return files


def reference_extras(seqs, cores):
    """the two other baselines BASELINE.md section 3 promises, with the reference's own code:
    (2) the split of one greedy iteration on two full-size sets (kmer_set_set.h:332-343),
    (4) kmerset-multiple-compress end to end on C1, bounded to its first C1_MERGES merges"""
    from _oracle import Ref, set_ref_seed
    if not Ref.available():
        return None, None
    ref = Ref()
    ref.lib.ref_set_log_level(4)
    tmp = tempfile.mkdtemp(prefix="kmsc_refx_")
    fj, fk = os.path.join(tmp, "j.txt"), os.path.join(tmp, "k.txt")
    write_spss(seqs[0], fj)
    write_spss(seqs[1], fk)
    sec, sizes = ref.split_stage(4, fj, fk, True, cores)
    keys = int(sizes[0] + sizes[1] + sizes[2] + sizes[3] + sizes[4])
    split = {"value": keys / sum(sec), "unit": "input+output keys/s", "seconds": sum(sec), "decode_s": sec[0], "intersection_s": sec[1],
             "sub_s": sec[2], "sizes": [int(x) for x in sizes],
             "what": f"ToKmerSet x2 + Intersection + Sub x2 on sets 0 and 1 (kmer_set_set.h:332-343), n_workers={cores}"}
    os.remove(fj)
    os.remove(fk)
    files = c1_files(tmp)
    set_ref_seed(77)
    t0 = time.time()
    out = ref.kmer_set_set(2, files, True, n_workers=cores, stop_after_weights=-C1_MERGES)
    dt = time.time() - t0
    c1 = {"seconds": dt, "merges": C1_MERGES, "unit": "s",
          "what": f"C1: reference KmerSetSet constructor on 8 sets x 1M 15-mers up to its merge {C1_MERGES + 1} (n_workers={cores}); "
                  "the GPU arm's `c1` runs the same bounded job through its executable"} if out["rc"] in (0, 1) else None
    for f in files:
        os.remove(f)
    os.rmdir(tmp)
    return split, c1


def ours_c1():
    """C1 through this repository's kmerset-multiple-compress, bounded like reference_extras"""
    host = ROOT / "kmer-sets-compression_b200" / "host"
    if subprocess.run(["make", "-s", "-C", str(host)], capture_output=True).returncode != 0:
        return None
    tmp = tempfile.mkdtemp(prefix="kmsc_c1_")
    files = c1_files(tmp)
    t0 = time.time()
    r = subprocess.run([str(host / "bin" / "kmerset-multiple-compress"), "--k=15", "--seed=77", f"--max_iterations={C1_MERGES}"] + files,
                       capture_output=True, text=True, timeout=900)
    wall = time.time() - t0
    for f in files:
        os.remove(f)
    os.rmdir(tmp)
    if r.returncode != 0:
        return {"error": r.stderr[-300:]}
    import re
    m = re.search(r"merges = (\d+), seconds = ([0-9.]+)", r.stderr)
    mi = re.search(r"device context ready, seconds = ([0-9.]+)", r.stderr)
    return {"seconds": float(m.group(2)) if m else None, "merges": int(m.group(1)) if m else None, "process_wall_s": wall, "unit": "s",
            "cuda_init_s": float(mi.group(1)) if mi else None,
            "what": f"C1: this repository's KmerSetSet constructor (device decode, weights, splits, device SPSS re-encode of the changed "
                    f"nodes) on 8 sets x 1M 15-mers, first {C1_MERGES} merges; cuda_init_s (the process's first CUDA call) is outside "
                    f"`seconds`, process_wall_s adds it and file loading"}


def reference_exact(seqs, steps, warmup, cores, n_total):
    """The GPU arm's job with the reference's code: all-bucket GetSampledKmerSet per set (reference
    lib/core/kmer_set_compact.h:120-203, one task per set as kmer_set_set.h:143-152) timed on ONE
    wave of min(cores, n) sets, then `steps` timed passes of the all-pairs two-pointer merge
    (kmer_set_set.h:158-219; the constructor's lambda cannot be handed all buckets, so the loop is
    the oracle's restatement of it, pthreads over the pairs) over those sets. Returns the phase
    rates and the projection to n_total sets."""
    import concurrent.futures as cf
    import synth
    from _oracle import Oracle, Ref
    o = Oracle()
    have_ref = Ref.available()
    ref = Ref() if have_ref else None
    n_w = min(cores, len(seqs))
    ids = np.arange(1 << N, dtype=np.int32)
    strings = [[p.decode() for p in synth.split_strings(s, K, 100000)] for s in seqs[:n_w]]

    def decode(i):
        if have_ref:
            offs, keys, _, _ = ref.sampled_set(4, strings[i], True, ids, n_workers=1)
        else:
            offs, keys = o.sampled_set(strings[i], K, N, True, ids)
        return offs, keys

    t0 = time.time()
    with cf.ThreadPoolExecutor(n_w) as ex:  # ctypes releases the GIL: one reference task per thread
        dec = list(ex.map(decode, range(n_w)))
    t_wave = time.time() - t0
    offs_l = [d[0] for d in dec]
    keys_l = [d[1].astype(np.uint32) for d in dec]   # the reference's KeyType for K = 23 (vector<uint32_t> buckets)
    times, visits = [], 0
    for it in range(warmup + steps):
        t0 = time.time()
        _, visits = o.pair_counts(offs_l, keys_l, KB, 1 << N, n_threads=cores)
        if it >= warmup:
            times.append(time.time() - t0)
    pair_rate = visits * len(times) / sum(times)
    keys_per_set = float(np.mean([int(o_[-1]) for o_ in offs_l]))
    visits_full = (n_total - 1) * keys_per_set * n_total
    waves = -(-n_total // cores)
    t_full = waves * t_wave + visits_full / pair_rate
    return {"kind": "reference" if have_ref else "port", "t_decode_wave_s": t_wave, "wave_sets": n_w, "pair_rate": pair_rate,
            "pair_step_s": float(np.mean(times)), "visits_step": int(visits), "visits_full": visits_full,
            "t_full_s": t_full, "value": visits_full / t_full, "waves": waves}


def reference_sampled(seqs, cores):
    """the reference's own constructor (lib/core/kmer_set_set.h:109-221) up to 'calculated initial
    weights' on ALL sets: GetSampledKmerSet over its 2 % bucket sample + all-pairs GetEdgeWeight"""
    from _oracle import Oracle, Ref, set_ref_seed
    if not Ref.available():
        return None
    set_ref_seed(4242)
    ref = Ref()
    ref.lib.ref_set_log_level(4)
    o = Oracle()
    tmp = tempfile.mkdtemp(prefix="kmsc_ref_")
    files = []
    for i, s in enumerate(seqs):
        f = os.path.join(tmp, f"{i}.txt")
        write_spss(s, f)
        files.append(f)
    c0 = ref.seed_counter()
    out = ref.kmer_set_set(4, files, True, n_workers=cores, stop_after_weights=True)
    for f in files:
        os.remove(f)
    os.rmdir(tmp)
    if out["rc"] != 1:
        return None
    ids = ref.random_ints(c0, (1 << N) // 50, 0, (1 << N) - 1)
    keys = sum(int(o.bucket_histogram(s, K, N)[ids].sum()) for s in seqs)
    v = (len(seqs) - 1) * keys
    t = out["phase_s"][0] + out["phase_s"][1]
    return {"value": v / t, "unit": UNIT, "seconds": t, "decode_s": out["phase_s"][0], "weights_s": out["phase_s"][1],
            "n_sets": len(seqs), "n_buckets": int(len(ids)), "key_visits": int(v),
            "what": "reference KmerSetSet constructor up to 'calculated initial weights' (its own 2% bucket sample), "
                    f"all {len(seqs)} sets, n_workers={cores}"}


def peaks_file():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


def run_c5(args, out):
    """BASELINE configs[4]: k = 15 dense-bitmap path, 512 sets, AND-popcount Gram (one GPU).
    Every set becomes a 2^30-bit bitmap (128 MiB); W = B B^T as tcgen05 kind::i8 MMAs on the 0/1-expanded
    bitmap words (csrc/bitmap.cu). The contraction is tensor-pipe bound, not HBM bound:
    both fractions are reported. Timing input: random k-mer sets (AND+POPC time does not depend on the
    density); parity: entries of W against the oracle's merge count."""
    import torch
    import kmsc
    from _oracle import Oracle
    K5, N5, KB5 = 15, 14, 2
    n = args.sets if args.sets > 0 else 512
    per_set = min(args.kmers, 2_000_000)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = kmsc.Context(0, stream.cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(777)
    base = torch.unique(torch.randint(0, 1 << 30, (per_set,), dtype=torch.int64, device=dev, generator=g))
    sets, host_sets = [], {}
    for i in range(n):
        # related sets: a random 90 % of a common pool plus a few private k-mers
        keep = torch.rand(base.numel(), device=dev, generator=g) < 0.9
        extra = torch.randint(0, 1 << 30, (per_set // 20,), dtype=torch.int64, device=dev, generator=g)
        km = torch.unique(torch.cat([base[keep], extra])).cpu().numpy().astype(np.uint64)
        sets.append(ctx.set_from_kmers(K5, N5, KB5, km))
        if i in (0, 1, n - 1):
            host_sets[i] = km
    sampler = ClockSampler(0)
    for _ in range(max(1, args.warmup)):
        W = ctx.bitmap_gram(sets)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        W = ctx.bitmap_gram(sets)          # bitmap fill + Gram + D2H of the matrix: the public call
    e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    o = Oracle()
    for i, j in ((0, 1), (0, n - 1), (1, n - 1)):
        want = int(o.merge_count(host_sets[i], host_sets[j]))
        assert int(W[i, j]) == want == int(W[j, i]), f"W[{i},{j}] = {int(W[i, j])}, oracle {want}"
    for i in host_sets:
        assert int(W[i, i]) == len(host_sets[i])
    bitmap_bytes = n * (1 << 27)
    blocks = -(-n // 256)
    macs = (blocks * (blocks + 1) // 2) * 256 * 256 * (1 << 30) if n > 128 else n * n * (1 << 30)
    word_ops = n * (n + 1) // 2 * (1 << 25)
    pk = peaks_file()
    hbm = float(pk.get("hbm_gbs", 6650.0))
    # int8 dense tensor peak: no measured entry; nominal 4.5 POP/s scaled by this pool's measured/nominal bf16 ratio
    int8_peak = 4500.0 * float(pk.get("bf16_tflops", 1590.0)) / 2250.0
    line = {"metric": "k=15 dense-bitmap Gram: AND-popcount 32-bit word-pairs per second (sets x sets x 2^25 words)",
            "value": word_ops / (ms / 1e3), "unit": "word-pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (0/1) x u8 -> s32 on tcgen05",
            "data": "synthetic", "config": {"workload": f"C5: {n} sets, k=15 bitmaps of 2^30 bits ({bitmap_bytes / 2**30:.0f} GiB), all-pairs AND-popcount Gram",
                                            "n_sets": n, "k": 15, "kmers_per_set": int(sets[0].n_keys),
                                            "l2": "64 GiB of bitmaps stream through the 126 MB L2; no flush needed"},
            "e2e": {"value": word_ops / (ms / 1e3), "unit": "word-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n * n * 8,
                    "note": "the sets are device-resident handles (the reference's KmerSet objects); the call builds the bitmaps, "
                            "runs the Gram and returns the matrix to the host"},
            "gpu_launches": int(ctx.launch_count() - launches0), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "bitmap_gram_tc_kernel", "achieved": 2 * macs / (ms / 1e3) / 1e12, "peak": int8_peak,
                         "unit": "TOP/s (int8)", "frac": 2 * macs / (ms / 1e3) / 1e12 / int8_peak, "traffic": None,
                         "peak_source": "nominal 4.5 POP/s dense int8 x (measured bf16 / nominal bf16) of MEASURED_PEAKS.json: no int8 entry is measured",
                         "hbm": {"achieved_gbs": bitmap_bytes / (ms / 1e3) / 1e9, "peak": hbm, "frac": bitmap_bytes / (ms / 1e3) / 1e9 / hbm,
                                 "algorithmic_bytes": bitmap_bytes},
                         "note": "time includes the bitmap fill and the D2H of the matrix (the whole public call)"},
            "oracle_check": {"entries": 6, "ok": True}}
    print(json.dumps(line), file=out, flush=True)


def gen_fasta_torch(n_bytes, read_len, genome_len, dev):
    """synthetic 2-line FASTA records on the GPU: reads of a random genome, both strands, 0.1 %
    substitutions, one N per 10 000 reads; fixed-width headers so that the file is a byte matrix"""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(2024)
    hdr = 10                                   # ">r" + 8 digits
    rec = hdr + 1 + read_len + 1
    n_reads = max(1, n_bytes // rec)
    genome = torch.randint(0, 4, (genome_len,), dtype=torch.uint8, device=dev, generator=g)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    out = torch.empty((n_reads, rec), dtype=torch.uint8, device=dev)
    step = 1 << 20
    ar = torch.arange(read_len, device=dev)
    for a in range(0, n_reads, step):
        b = min(n_reads, a + step)
        m = b - a
        start = torch.randint(0, genome_len - read_len, (m,), device=dev, generator=g)
        codes = genome[start[:, None] + ar[None, :]]
        rcm = torch.rand(m, device=dev, generator=g) < 0.5
        codes = torch.where(rcm[:, None], (3 - codes).flip(1), codes)
        err = torch.rand((m, read_len), device=dev, generator=g) < 0.001
        codes = torch.where(err, (codes + torch.randint(1, 4, (m, read_len), dtype=torch.uint8, device=dev, generator=g)) & 3, codes)
        seq = lut[codes.long()]
        hasn = torch.rand(m, device=dev, generator=g) < 1e-4
        pos = torch.randint(0, read_len, (m,), device=dev, generator=g)
        seq[hasn.nonzero().flatten(), pos[hasn]] = 78   # 'N'
        ids = torch.arange(a, b, device=dev)
        out[a:b, 0] = 62
        out[a:b, 1] = 114
        for d in range(8):
            out[a:b, 2 + d] = (48 + (ids // (10 ** (7 - d))) % 10).to(torch.uint8)
        out[a:b, hdr] = 10
        out[a:b, hdr + 1:hdr + 1 + read_len] = seq
        out[a:b, rec - 1] = 10
    return out, rec


def run_c4(args, out):
    """BASELINE configs[3]: kmerset-build counting on synthetic FASTA reads, k = 31, cutoff 4
    (KmerCounter::FromFASTA + ToKmerSet, reference lib/core/kmer_counter.h:64-243). A slice of the 20 GB
    file (--bytes) streams from pinned host memory through the chunked counter (kmsc_counter_*);
    metric = bases counted per second end to end. The reference's own KmerCounter runs beside it on a
    bounded slice with all host threads."""
    import torch
    import kmsc
    K4, N4, KB4 = 31, 14, 8
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = kmsc.Context(0, stream.cuda_stream)
    read_len = 150
    genome_len = 20_000_000
    recs, rec = gen_fasta_torch(args.bytes, read_len, genome_len, dev)
    n_reads = recs.shape[0]
    host = torch.empty(recs.numel(), dtype=torch.uint8, pin_memory=True)
    host.copy_(recs.flatten())
    del recs
    torch.cuda.synchronize()
    data = host.numpy()
    per_chunk = max(1, args.chunk_bytes // rec) * rec
    chunks = [data[a:a + per_chunk] for a in range(0, len(data), per_chunk)]
    bases = n_reads * read_len

    def step():
        s, cut, nd = ctx.count_chunks(K4, N4, KB4, chunks, canonical=True, cutoff=4, fasta=True)
        r = (s.n_keys, cut, nd)
        s.free()
        return r

    sampler = ClockSampler(0)
    for _ in range(max(1, min(args.warmup, 2))):
        res = step()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = ctx.launch_count()
    steps = max(1, min(args.steps, 5))
    t0 = time.time()
    for _ in range(steps):
        res = step()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / steps
    clocks = sampler.stop()
    # parity on a slice against the oracle (FromReads semantics) and the reference's own counter beside it
    from _oracle import Oracle, Ref
    o = Oracle()
    n_chk = min(n_reads, 20000)
    reads_chk = [bytes(data[i * rec + 11:i * rec + 11 + read_len]).decode() for i in range(n_chk)]
    km, cnt = o.count_reads(reads_chk, K4, True)
    kept, cut_o = o.counter_to_set(km, cnt, 2)
    s2, cut2, nd2 = ctx.count_fasta(K4, N4, KB4, bytes(data[:n_chk * rec]), canonical=True, cutoff=2)
    assert (s2.n_keys, cut2, nd2) == (len(kept), cut_o, len(km)), "counting differs from the oracle on the check slice"
    assert s2.Hash() == o.set_hash(kept)
    cpu = None
    if Ref.available() and not args.no_cpu_baseline:
        ref = Ref()
        n_ref = min(n_reads, 400_000)
        reads_ref = [bytes(data[i * rec + 11:i * rec + 11 + read_len]).decode() for i in range(n_ref)]
        t1 = time.time()
        ref.count_reads(5, reads_ref, True, 4, n_workers=os.cpu_count() or 1)
        t_ref = time.time() - t1
        cpu = {"value": n_ref * read_len / t_ref, "unit": "bases/s", "cores": os.cpu_count() or 1, "kind": "reference",
               "sample": f"KmerCounter::FromReads + ToKmerSet(4) (+ the test driver's read-back of the counts) on the first {n_ref} reads "
                         f"({n_ref * read_len / 1e6:.0f} Mbases), n_workers = all cores; {t_ref:.1f} s"}
    pk = peaks_file()
    hbm = float(pk.get("hbm_gbs", 6650.0))
    line = {"metric": "k-mer counting throughput, FASTA bytes in host memory -> counted set (k=31, cutoff 4)", "value": bases / dt,
            "unit": "bases/s", "n_gpus": 1, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"C4 slice: {len(data) / 1e9:.2f} GB of 2-line FASTA ({n_reads} reads x {read_len} bp, both strands, 0.1% "
                                   f"substitutions, genome {genome_len} bp), canonical 31-mers <31,14,uint64>, cutoff 4, streamed in "
                                   f"{len(chunks)} chunks of whole records", "k": 31, "cutoff": 4, "bytes": int(len(data)),
                       "l2": "every chunk (256 MB) exceeds the 126 MB L2"},
            "e2e": {"value": bases / dt, "unit": "bases/s", "h2d_bytes_per_step": int(len(data)), "d2h_bytes_per_step": 64,
                    "note": "value IS the end-to-end number: the bytes start in pinned host memory"},
            "gpu_launches": int(ctx.launch_count() - launches0), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "P1 pipeline (classify/pack + partition sort + run-length count + merge)",
                         "achieved": len(data) / dt / 1e9, "peak": hbm, "unit": "GB/s", "frac": len(data) / dt / 1e9 / hbm, "traffic": None,
                         "note": "algorithmic bytes = the file read once; the step is bound by the per-chunk sort passes and the host-to-device copy"},
            "result": {"kept": int(res[0]), "cutoff_count": int(res[1]), "distinct": int(res[2])},
            "cpu_baseline": cpu, "oracle_check": {"reads": n_chk, "ok": True}}
    print(json.dumps(line), file=out, flush=True)


def main():
    args = parse_args()
    strong = args.workload == "C3"
    if args.sets <= 0:
        args.sets = {"C3": 256, "C5": 512}.get(args.workload, 64)
    # stdout carries exactly ONE line (the JSON): anything a library prints there (NCCL's version
    # banner, for one) goes to stderr instead
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    G = args.kmers + K - 1  # bases per set per GPU
    cores = os.cpu_count() or 1
    config = {"workload": (f"C3: {args.sets} sets x {args.kmers} canonical {K}-mers in all (<{K},{N},uint32>), "
                           if strong else
                           f"C2: {args.sets} sets x {args.kmers} canonical {K}-mers per GPU (<{K},{N},uint32>), ")
                          + f"binary phylogeny p={args.p}, all-pairs intersection matrix over all {1 << N} buckets (exact)",
              "n_sets": args.sets, ("kmers_per_set" if strong else "kmers_per_set_per_gpu"): args.kmers, "k": K, "bucket_bits": N,
              "parallelism": f"prefix-sharded x{world}, all-reduce inside kmsc_pair_counts" if world > 1 else "single GPU",
              "l2": f"inputs ({args.sets * args.kmers * 4 / (1 if not strong else world) / 1e9:.2f} GB/GPU) exceed the 126 MB L2; no flush needed"}

    if args.workload in ("C4", "C5") and args.impl != "reference":
        if rank == 0:
            (run_c4 if args.workload == "C4" else run_c5)(args, out)
        return

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        seqs, how = gen_codes(args.sets, G, args.p)
        ex = reference_exact(seqs, args.steps, args.warmup, cores, args.sets)
        # N ranks: the GPU arm's job holds `world` times the k-mers per set (weak scaling); both
        # reference phases are linear in the sequence length, so the measured rates are scaled
        visits_full = ex["visits_full"] * world
        t_full = ex["waves"] * ex["t_decode_wave_s"] * world + visits_full / ex["pair_rate"]
        val = visits_full / t_full
        smp = None if args.no_ref_sampled else reference_sampled(seqs, cores)
        split_ref, c1_ref = (None, None) if args.no_ref_sampled else reference_extras(seqs, cores)
        sample = (f"exact job of the config with the reference's code: GetSampledKmerSet over all {1 << N} buckets timed on one "
                  f"wave of {ex['wave_sets']} of the {args.sets} sets ({ex['t_decode_wave_s']:.2f} s, one task per set), then each step = "
                  f"the all-pairs two-pointer merge over those {ex['wave_sets']} sets on {cores} threads ({ex['pair_step_s']:.3f} s, "
                  f"{ex['pair_rate']:.3e} key-visits/s); value = key-visits of the whole {args.sets}-set job / "
                  f"({ex['waves']} decode waves + all pairs at that rate) = {t_full:.1f} s per job"
                  + (f", sequence length scaled x{world}" if world > 1 else "") + f"; data: {how}")
        cfg = dict(config)
        cfg["reference_sample"] = sample
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg, "impl": "reference",
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": ex["kind"], "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "phases": {k: ex[k] for k in ("t_decode_wave_s", "wave_sets", "waves", "pair_rate", "pair_step_s", "visits_step")},
                "sampled": smp, "split_stage": split_ref, "c1": c1_ref, "gpu_launches": 0}
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
    if world > 1:
        # the library owns the communicator: the id travels through torch.distributed once
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(kmsc.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))

    # data: genome grows with the number of ranks; every rank builds the same sequences
    shard_rank, shard_world = rank, world
    if args.emulate and world == 1:
        shard_rank, shard_world = (int(x) for x in args.emulate.split("/"))
    Gtot = args.kmers * (1 if strong else shard_world) + K - 1
    seqs = gen_sequences_torch(args.sets, Gtot, args.p, dev)
    str_offs = np.array([0, Gtot], np.int64)
    # CPU copies for the checks outside the timed regions: the oracle check of W and the CPU baseline
    check_ids = sorted({0, 1, 2, args.sets - 2, args.sets - 1} & set(range(args.sets)))
    keep = set(check_ids) | (set(range(min(cores, args.sets))) if world == 1 and not args.no_cpu_baseline else set())
    codes_cpu = {i: seqs[i].cpu().numpy() for i in sorted(keep)} if world == 1 else {}
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
        ctx.pair_counts_device(sets, d_out.data_ptr())   # N > 1: the partial matrices are all-reduced inside

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

    # ---- parity at full size, outside the timed region: entries of W against the oracle's two-pointer
    # merge (reference kmer_set_set.h:158-184) over k-mer sets extracted on the CPU with numpy from the
    # same sequences. A mismatch fails the run.
    oracle_check = None
    if world == 1 and not args.no_check and len(check_ids) >= 2:
        import synth
        from _oracle import Oracle
        o = Oracle()
        ksets = {i: synth.kmer_set_of(codes_cpu[i], K, True) for i in check_ids}
        checked = []
        for i in check_ids:
            assert int(W[i, i]) == len(ksets[i]), f"W[{i},{i}] = {int(W[i, i])} but the set holds {len(ksets[i])} k-mers"
            checked.append([i, i, int(W[i, i])])
        for a in range(len(check_ids)):
            for b in range(a + 1, len(check_ids)):
                i, j = check_ids[a], check_ids[b]
                want = int(o.merge_count(ksets[i], ksets[j]))
                assert int(W[i, j]) == want and int(W[j, i]) == want, f"W[{i},{j}] = {int(W[i, j])}, oracle {want}"
                checked.append([i, j, want])
        oracle_check = {"entries": len(checked), "ok": True, "W_0_last": int(W[0, n - 1]),
                        "how": "numpy k-mer sets of the same sequences + oracle merge_count; every listed entry equal"}
        del ksets

    # ---- "pairwise-weight plus diff" stage: the matrix + one split per spanning-tree edge ----
    # (north_star: MST over d(i,j) = |Si| + |Sj| - 2 W[i][j]; per edge the two difference sets
    # via kmsc_pair_split). Bytes are the algorithmic B_w + sum B_s of SURVEY 8(d).
    stage = None
    split_pair = None
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
            ctx.pair_counts_device(sets, d_loc.data_ptr(), partial=True)
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
        # one greedy-iteration split (n, j \\ n, k \\ n) of sets 0 and 1: the reference arm's `split_stage`
        if world == 1:
            for _ in range(2):
                for x in ctx.pair_split(sets[0], sets[1]):
                    x.free()
            barrier()
            es0.record(stream)
            for _ in range(5):
                outs3 = ctx.pair_split(sets[0], sets[1])
                k3 = sets[0].n_keys + sets[1].n_keys + sum(x.n_keys for x in outs3)
                for x in outs3:
                    x.free()
            es1.record(stream)
            barrier()
            pair_ms = es0.elapsed_time(es1) / 5
            split_pair = {"value": k3 / (pair_ms / 1e3), "unit": "input+output keys/s", "ms": pair_ms,
                          "what": "kmsc_pair_split of sets 0 and 1, all three outputs, sets resident on the device"}
        stage = {"what": "all-pairs matrix + the two difference sets of each of the n-1 MST edges (one kmsc_pair_split_batch), per GPU",
                 "ms": w_ms + split_ms, "weights_ms": w_ms, "splits_ms": split_ms, "n_splits": len(edges),
                 "algorithmic_bytes": w_bytes + split_bytes,
                 "achieved_gbs": (w_bytes + split_bytes) / ((w_ms + split_ms) / 1e3) / 1e9}

    # ---- f1: SPSS construction of one full-size set on the device (kmsc_spss_build), next to the reference's
    # GetSPSSCanonical (lib/core/spss.h:1039-1858) on a bounded slice of the same sequence ----
    spss = None
    if world == 1 and not args.no_stage:
        try:
            ctx.spss_build(sets[0], True, 0, fetch=False)   # allocations
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_str, n_chr = ctx.spss_build(sets[0], True, 0, fetch=False)
            ctx.sync()
            dt = time.perf_counter() - t0
            n_km = int(W[0, 0])
            spss = {"value": n_km / dt, "unit": "k-mers/s", "ms": dt * 1e3, "kmers": n_km, "strings": int(n_str), "chars": int(n_chr),
                    "what": "kmsc_spss_build on set 0 (canonical, 8 matching rounds), text left on the device"}
            if not args.no_cpu_baseline:
                from _oracle import Ref, CFG_BY_K
                if Ref.available():
                    import synth
                    sl = synth.kmer_set_of(codes_cpu[0][:500_000], K, True)
                    ref = Ref()
                    t0 = time.perf_counter()
                    theirs, _w = ref.spss_from_set(CFG_BY_K[K], sl, True, fast=True, n_workers=cores)
                    dt_r = time.perf_counter() - t0
                    spss["reference"] = {"value": len(sl) / dt_r, "unit": "k-mers/s", "seconds": dt_r, "kmers": int(len(sl)),
                                         "strings": len(theirs), "chars": int(sum(map(len, theirs))), "cores": cores,
                                         "what": "the reference's GetSPSSCanonical(fast) on the k-mers of the first 500 000 bases of "
                                                 "the same sequence, n_workers = all cores (includes the test driver's KmerSet build)"}
        except Exception as ex:   # a diagnostic: it must never take the bench line down
            spss = {"error": str(ex)[:200]}

    # ---- e2e: host packed SPSS -> device CSR -> matrix -> host ------------------------
    e2e = None
    sampled = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 5))
        host_out = np.zeros((n, n), np.int64)

        def step_e2e_single():
            ss = build_sets()
            ctx.pair_counts_device(ss, d_out.data_ptr())
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

        phase_ms, phase_calls = {}, {}

        def timed(name, fn):
            if not args.profile_e2e:
                return fn()
            phase_calls[name] = phase_calls.get(name, 0) + 1
            if phase_calls[name] <= 2:   # connection set-up, allocation pools
                return fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            phase_ms[name] = phase_ms.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
            return r

        def step_e2e_exchange():
            mine = [rank + j * world for j in range(m_own)]
            full = timed("decode", lambda: ctx.sets_from_packed_batch(K, N, KB, None, [str_offs] * m_own,
                                                                    words_ptrs=[pinned[i].data_ptr() for i in mine]))
            ss = timed("exchange", lambda: ctx.sets_exchange(full, cuts_np, n))     # one grouped NCCL exchange inside the library
            timed("pair_counts", lambda: ctx.pair_counts_device(ss, d_out.data_ptr()))
            host_out[:] = timed("d2h", lambda: d_out.cpu().numpy().reshape(n, n))
            for s in ss + full:
                s.free()

        # The same step with the exchange overlapped with the decode: a second context (its own stream, no
        # communicator) decodes, the first one exchanges.
        ctx2 = None
        if m_own >= 2 and not args.no_e2e_overlap:
            stream2 = torch.cuda.Stream(device=dev)
            ctx2 = kmsc.Context(local_rank, stream2.cuda_stream)

        def step_e2e_overlap():
            # P pieces of the rank's sets: piece p is exchanged (thread, first context) while piece p + 1 is decoded
            # (second context); a piece is an independent collection of len(piece) * world sets, and the outputs of
            # the pieces concatenate to the global set order (global id = rank + j * world)
            P = min(args.e2e_pieces, m_own)
            mine = [rank + j * world for j in range(m_own)]
            bounds = [m_own * q // P for q in range(P + 1)]
            outs, fulls, th = [None] * P, [], None

            def ex(q, full):
                outs[q] = ctx.sets_exchange(full, cuts_np, len(full) * world)
            for q in range(P):
                ids = mine[bounds[q]:bounds[q + 1]]
                full = ctx2.sets_from_packed_batch(K, N, KB, None, [str_offs] * len(ids), words_ptrs=[pinned[i].data_ptr() for i in ids])
                ctx2.sync()
                fulls += full
                if th is not None:
                    th.join()
                th = threading.Thread(target=ex, args=(q, full))
                th.start()
            th.join()
            ss = [s_ for o in outs for s_ in o]
            ctx.pair_counts_device(ss, d_out.data_ptr())
            host_out[:] = d_out.cpu().numpy().reshape(n, n)
            for s_ in ss + fulls:
                s_.free()

        step_e2e = (step_e2e_overlap if ctx2 is not None and not args.profile_e2e else step_e2e_exchange) if m_own > 0 else step_e2e_single

        for s in sets:
            s.free()
        step_e2e()
        if args.profile_e2e:
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
        if args.profile_e2e and phase_ms:
            print(f"[rank {rank}] e2e phases, ms per step: " + ", ".join(f"{k} {v / max(1, phase_calls[k] - 2):.2f}" for k, v in phase_ms.items()),
                  file=sys.stderr, flush=True)
        assert np.array_equal(host_out, W), "e2e matrix differs from the resident run"
        # the reference's own mode: weights over a 2 % bucket sample (kmer_set_set.h:123-124), same sets
        if world == 1:
            ids = np.sort(np.random.default_rng(4242).choice(1 << N, (1 << N) // 50, replace=False)).astype(np.int32)

            def step_sampled():
                ss = build_sets()
                ctx.pair_counts_device(ss, d_out.data_ptr(), bucket_ids=ids)
                host_out[:] = d_out.cpu().numpy().reshape(n, n)
                for s in ss:
                    s.free()
            step_sampled()
            barrier()
            e0.record(stream)
            for _ in range(e2e_steps):
                step_sampled()
            e1.record(stream)
            barrier()
            ms_s = e0.elapsed_time(e1) / e2e_steps
            v_s = (n - 1) * int(np.trace(host_out))
            sampled = {"value": v_s / (ms_s / 1e3), "unit": UNIT, "ms_per_step": ms_s, "n_buckets": int(len(ids)), "key_visits": v_s,
                       "what": "host packed SPSS -> device sets -> matrix over a 2% bucket sample -> host (the reference "
                               "constructor's own job up to 'calculated initial weights'; compare with the reference arm's `sampled`)"}
        e2e = {"value": visits_total * e2e_steps / (ms_e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(nbytes_in // world) if m_own > 0 else int(nbytes_in),
               "how": ("decode split by set over the ranks (kmsc_sets_from_packed_batch) + kmsc_sets_exchange (one grouped NCCL send/recv of "
                       "the prefix slices) + kmsc_pair_counts_device (all-reduce inside)" + (f"; the rank's sets go in {min(args.e2e_pieces, m_own)} pieces: a piece is "
                       "exchanged while a second context decodes the next one" if ctx2 is not None and not args.profile_e2e else "")
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
    kernel_name = {1: "pair_counts_stream_kernel", 3: "pair_counts_lane_kernel"}.get(p3_build, "pair_counts_kernel")
    try:
        tr = json.loads((ROOT / "profiles" / "p3_traffic.json").read_text())["kernels"][kernel_name]
        if args.sets == 64 and args.kmers == 10_000_000 and K == 23:
            traffic = int(tr["dram_bytes_read"]) + int(tr["dram_bytes_write"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_name,
                "build": {1: "warp-wide multiway merge (related sets)", 3: "lane-private tables (one fine bucket per lane)"}.get(p3_build, "shared-memory hash table"),
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": algo_bytes / max(1, main_launches),
                "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": main_ms / ms if ms else None,
                "plan_ms_per_step": plan_ms / args.steps,
                "note": "the build is chosen from the redundancy rho (keys per distinct key) of the sets; the library measures "
                        "rho on 1/64 of the buckets on the FIRST call of a shape (one extra hash-build pass over that slice) "
                        "and remembers it in the context: the warm-up steps pay that probe, the timed steps do not"}

    if stage is not None:
        stage["frac_of_hbm_peak"] = stage["achieved_gbs"] / peak
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            n_w = min(cores, args.sets)
            ex = reference_exact([codes_cpu[i] for i in range(n_w)], 2, 0, cores, args.sets)
            cpu_baseline = {"value": ex["value"], "unit": UNIT, "cores": cores, "kind": ex["kind"],
                            "sample": f"the reference's code on {n_w} of the {args.sets} sets: all-bucket GetSampledKmerSet, one task per set "
                                      f"({ex['t_decode_wave_s']:.1f} s per wave), 2 passes of the all-pairs two-pointer merge over them "
                                      f"({ex['pair_rate']:.3e} key-visits/s on {cores} threads); value = the whole {args.sets}-set job at "
                                      f"those rates ({ex['t_full_s']:.1f} s). bench.py --impl reference is the full arm"}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {ex}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stage": stage, "cpu_baseline": cpu_baseline,
            "sampled": sampled, "split_stage": split_pair, "spss": spss, "c1": (ours_c1() if world == 1 and not args.no_c1 else None),
            "oracle_check": oracle_check,
            "check": {"W01": int(W[0, 1]), "W_diag0": int(W[0, 0]), "keys_per_gpu": int(keys_local),
                      "p3_stats": ctx.pair_counts_stats() if args.no_e2e and args.no_stage else None, "buckets": [int(lo), int(hi)]}}
    if per_rank is not None:
        line["per_rank"] = per_rank
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
