"""ctypes bindings for the TEST-ONLY CPU checkers under oracle/.

* ``Oracle``  -> oracle/libkmsc_oracle.so   (plain-C restatement, kmsc_oracle.c)
* ``Ref``     -> oracle/_ref/libkmsc_ref.so (the reference's unmodified headers,
  compiled by oracle/Makefile; only where it was prebuilt)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "libkmsc_oracle.so"
REF_SO = ORACLE_DIR / "_ref" / "libkmsc_ref.so"

# config id -> (K, N, key_bytes); mirrors oracle/ref_driver.cc
CONFIGS = {0: (5, 3, 1), 1: (9, 10, 1), 2: (15, 14, 2), 3: (19, 10, 4), 4: (23, 14, 4), 5: (31, 14, 8)}
CFG_BY_K = {k: cfg for cfg, (k, _, _) in CONFIGS.items()}

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


def build_oracle() -> None:
    """(Re)build the checkers; a no-op for _ref where /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "all"], check=True)


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def _strs(strings):
    arr = (C.c_char_p * max(1, len(strings)))()
    for i, s in enumerate(strings):
        arr[i] = s.encode() if isinstance(s, str) else s
    return arr


class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            build_oracle()
        L = self.lib = C.CDLL(str(ORACLE_SO))
        L.kmsc_o_kmer_from_string.argtypes = [C.c_char_p, C.c_int, u64p]
        L.kmsc_o_kmer_to_string.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        for f in ("kmsc_o_complement", "kmsc_o_canonical"):
            getattr(L, f).argtypes = [C.c_uint64, C.c_int]
            getattr(L, f).restype = C.c_uint64
        for f in ("kmsc_o_next", "kmsc_o_prev"):
            getattr(L, f).argtypes = [C.c_uint64, C.c_int, C.c_char]
            getattr(L, f).restype = C.c_uint64
        L.kmsc_o_bucket_key.argtypes = [C.c_uint64, C.c_int, C.c_int, i32p, u64p]
        L.kmsc_o_from_bucket_key.argtypes = [C.c_int32, C.c_uint64, C.c_int, C.c_int]
        L.kmsc_o_from_bucket_key.restype = C.c_uint64
        L.kmsc_o_add_with_max_u8.argtypes = [C.c_uint8, C.c_uint8]
        L.kmsc_o_add_with_max_u8.restype = C.c_uint8
        L.kmsc_o_fasta_validate.argtypes = [C.POINTER(C.c_char_p), C.c_int64]
        L.kmsc_o_count_reads.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int,
                                         C.POINTER(u64p), C.POINTER(u8p)]
        L.kmsc_o_count_reads.restype = C.c_int64
        L.kmsc_o_counter_to_set.argtypes = [u64p, u8p, C.c_int64, C.c_uint8, u64p, i64p]
        L.kmsc_o_counter_to_set.restype = C.c_int64
        L.kmsc_o_compact_pack.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_int, u64p, u32p]
        L.kmsc_o_compact_pack.restype = C.c_int64
        L.kmsc_o_compact_unpack.argtypes = [u64p, C.c_int64, C.c_int64, C.c_char_p]
        L.kmsc_o_compact_size.argtypes = [u32p, C.c_int64]
        L.kmsc_o_compact_size.restype = C.c_int64
        L.kmsc_o_compact_weight.argtypes = [u32p, C.c_int64, C.c_int]
        L.kmsc_o_compact_weight.restype = C.c_int64
        L.kmsc_o_spss_kmers.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, u64p]
        L.kmsc_o_spss_kmers.restype = C.c_int64
        L.kmsc_o_sampled_set.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, C.c_int,
                                         i32p, C.c_int32, i64p, u64p]
        L.kmsc_o_sampled_set.restype = C.c_int64
        L.kmsc_o_set_from_spss.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, u64p]
        L.kmsc_o_set_from_spss.restype = C.c_int64
        for f in ("kmsc_o_set_add", "kmsc_o_set_sub", "kmsc_o_set_intersection"):
            getattr(L, f).argtypes = [u64p, C.c_int64, u64p, C.c_int64, u64p]
            getattr(L, f).restype = C.c_int64
        L.kmsc_o_set_diff.argtypes = [u64p, C.c_int64, u64p, C.c_int64]
        L.kmsc_o_set_diff.restype = C.c_int64
        L.kmsc_o_set_hash.argtypes = [u64p, C.c_int64]
        L.kmsc_o_set_hash.restype = C.c_uint64
        L.kmsc_o_bucket_offsets.argtypes = [u64p, C.c_int64, C.c_int, C.c_int, i64p]
        L.kmsc_o_merge_count.argtypes = [u64p, C.c_int64, u64p, C.c_int64]
        L.kmsc_o_merge_count.restype = C.c_int64
        L.kmsc_o_pair_counts.argtypes = [C.POINTER(i64p), C.POINTER(C.c_void_p), C.c_int32, C.c_int,
                                         i32p, C.c_int32, C.c_int32, C.c_int, i64p, i64p]
        L.kmsc_o_greedy_interval.argtypes = [C.c_int32]
        L.kmsc_o_greedy_threshold.argtypes = [C.c_int32]
        L.kmsc_o_greedy_threshold.restype = C.c_float
        L.kmsc_o_greedy_should_stop.argtypes = [C.c_int64, C.c_int64, C.c_int32]
        L.kmsc_o_greedy_argmax.argtypes = [i64p, C.c_int32, i32p, i32p]
        L.kmsc_o_greedy_argmax.restype = C.c_int64
        L.kmsc_o_bucket_histogram.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, i64p]
        L.kmsc_o_bucket_histogram.restype = None
        L.kmsc_o_mst.argtypes = [i64p, C.c_int32, i32p, i64p]
        L.kmsc_o_mst.restype = C.c_int32
        L.kmsc_o_dsu_new.argtypes = [C.c_int32]
        L.kmsc_o_dsu_new.restype = C.c_void_p
        L.kmsc_o_dsu_free.argtypes = [C.c_void_p]
        L.kmsc_o_dsu_find.argtypes = [C.c_void_p, C.c_int32]
        L.kmsc_o_dsu_same.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.kmsc_o_dsu_unite.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.kmsc_o_svb0124_max_bytes.argtypes = [C.c_uint32]
        L.kmsc_o_svb0124_max_bytes.restype = C.c_size_t
        L.kmsc_o_svb0124_encode.argtypes = [u32p, C.c_uint32, u8p]
        L.kmsc_o_svb0124_encode.restype = C.c_size_t
        L.kmsc_o_svb0124_decode.argtypes = [u8p, u32p, C.c_uint32]
        L.kmsc_o_svb0124_decode.restype = C.c_size_t
        L.kmsc_o_codec_encode.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64)]
        L.kmsc_o_codec_encode.restype = C.c_void_p
        L.kmsc_o_codec_decode.argtypes = [u8p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
        L.kmsc_o_codec_decode.restype = C.c_int
        L.kmsc_o_free.argtypes = [C.c_void_p]

    # -- Kmer ---------------------------------------------------------------
    def bits(self, s: str) -> int:
        out = C.c_uint64()
        if self.lib.kmsc_o_kmer_from_string(s.encode(), len(s), C.byref(out)) != 0:
            raise ValueError(s)
        return out.value

    def string(self, bits: int, K: int) -> str:
        buf = C.create_string_buffer(K + 1)
        self.lib.kmsc_o_kmer_to_string(bits, K, buf)
        return buf.value.decode()

    def complement(self, bits, K): return self.lib.kmsc_o_complement(bits, K)
    def canonical(self, bits, K): return self.lib.kmsc_o_canonical(bits, K)
    def next(self, bits, K, c): return self.lib.kmsc_o_next(bits, K, c.encode())
    def prev(self, bits, K, c): return self.lib.kmsc_o_prev(bits, K, c.encode())

    def bucket_key(self, bits, K, N):
        b, k = C.c_int32(), C.c_uint64()
        self.lib.kmsc_o_bucket_key(bits, K, N, C.byref(b), C.byref(k))
        return b.value, k.value

    def from_bucket_key(self, b, k, K, N): return self.lib.kmsc_o_from_bucket_key(b, k, K, N)

    # -- KmerCounter ----------------------------------------------------------
    def fasta_validate(self, lines): return self.lib.kmsc_o_fasta_validate(_strs(lines), len(lines))

    def count_reads(self, reads, K, canonical):
        kp, cp = u64p(), u8p()
        n = self.lib.kmsc_o_count_reads(_strs(reads), len(reads), K, int(canonical), C.byref(kp), C.byref(cp))
        if n < 0:
            raise ValueError("bad read")
        kmers = np.ctypeslib.as_array(kp, shape=(max(n, 1),))[:n].copy()
        counts = np.ctypeslib.as_array(cp, shape=(max(n, 1),))[:n].copy()
        self.lib.kmsc_o_free(kp)
        self.lib.kmsc_o_free(cp)
        return kmers, counts

    def counter_to_set(self, kmers, counts, cutoff):
        kept = np.empty(len(kmers), np.uint64)
        cut = C.c_int64()
        kmers = np.ascontiguousarray(kmers, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint8)
        m = self.lib.kmsc_o_counter_to_set(_ptr(kmers, u64p), _ptr(counts, u8p), len(kmers), cutoff,
                                           _ptr(kept, u64p), C.byref(cut))
        return kept[:m].copy(), cut.value

    # -- KmerSetCompact -------------------------------------------------------
    def compact_pack(self, strings, K):
        total = sum(len(s) for s in strings)
        words = np.zeros((total + 31) // 32 + 1, np.uint64)
        lens = np.zeros(max(1, len(strings)), np.uint32)
        n = self.lib.kmsc_o_compact_pack(_strs(strings), len(strings), K, _ptr(words, u64p), _ptr(lens, u32p))
        assert n == total
        return words, lens[:len(strings)]

    def compact_unpack(self, words, lens, K):
        out, pos = [], 0
        for l in lens:
            n = int(l) + K
            buf = C.create_string_buffer(n + 1)
            self.lib.kmsc_o_compact_unpack(_ptr(words, u64p), pos, n, buf)
            out.append(buf.value.decode())
            pos += n
        return out

    def compact_size(self, lens):
        lens = np.ascontiguousarray(lens, np.uint32)
        return self.lib.kmsc_o_compact_size(_ptr(lens, u32p), len(lens))

    def compact_weight(self, lens, K):
        lens = np.ascontiguousarray(lens, np.uint32)
        return self.lib.kmsc_o_compact_weight(_ptr(lens, u32p), len(lens), K)

    def _npos(self, strings, K): return sum(max(0, len(s) - K + 1) for s in strings)

    def spss_kmers(self, strings, K, canonical):
        out = np.empty(max(1, self._npos(strings, K)), np.uint64)
        n = self.lib.kmsc_o_spss_kmers(_strs(strings), len(strings), K, int(canonical), _ptr(out, u64p))
        return out[:n].copy()

    def sampled_set(self, strings, K, N, canonical, bucket_ids):
        ids = np.ascontiguousarray(bucket_ids, np.int32)
        offs = np.zeros(len(ids) + 1, np.int64)
        keys = np.empty(max(1, self._npos(strings, K)), np.uint64)
        n = self.lib.kmsc_o_sampled_set(_strs(strings), len(strings), K, N, int(canonical), _ptr(ids, i32p),
                                        len(ids), _ptr(offs, i64p), _ptr(keys, u64p))
        return offs, keys[:n].copy()

    def set_from_spss(self, strings, K, canonical):
        out = np.empty(max(1, self._npos(strings, K)), np.uint64)
        n = self.lib.kmsc_o_set_from_spss(_strs(strings), len(strings), K, int(canonical), _ptr(out, u64p))
        return out[:n].copy()

    # -- KmerSet algebra --------------------------------------------------------
    def _setop(self, f, a, b, cap):
        a = np.ascontiguousarray(a, np.uint64)
        b = np.ascontiguousarray(b, np.uint64)
        out = np.empty(max(1, cap), np.uint64)
        n = f(_ptr(a, u64p), len(a), _ptr(b, u64p), len(b), _ptr(out, u64p))
        return out[:n].copy()

    def set_add(self, a, b): return self._setop(self.lib.kmsc_o_set_add, a, b, len(a) + len(b))
    def set_sub(self, a, b): return self._setop(self.lib.kmsc_o_set_sub, a, b, len(a))
    def set_intersection(self, a, b): return self._setop(self.lib.kmsc_o_set_intersection, a, b, len(a))

    def set_diff(self, a, b):
        a = np.ascontiguousarray(a, np.uint64)
        b = np.ascontiguousarray(b, np.uint64)
        return self.lib.kmsc_o_set_diff(_ptr(a, u64p), len(a), _ptr(b, u64p), len(b))

    def set_hash(self, a):
        a = np.ascontiguousarray(a, np.uint64)
        return self.lib.kmsc_o_set_hash(_ptr(a, u64p), len(a))

    def bucket_offsets(self, kmers, K, N):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        offs = np.zeros((1 << N) + 1, np.int64)
        self.lib.kmsc_o_bucket_offsets(_ptr(kmers, u64p), len(kmers), K, N, _ptr(offs, i64p))
        return offs

    def to_csr(self, kmers, K, N, key_bytes):
        """ascending distinct k-mers -> (offs int64[2^N+1], keys of key_bytes)."""
        kmers = np.ascontiguousarray(kmers, np.uint64)
        offs = self.bucket_offsets(kmers, K, N)
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[key_bytes]
        keys = (kmers & np.uint64((1 << (2 * K - N)) - 1)).astype(dt)
        return offs, keys

    # -- GetEdgeWeight ------------------------------------------------------------
    def merge_count(self, a, b):
        a = np.ascontiguousarray(a, np.uint64)
        b = np.ascontiguousarray(b, np.uint64)
        return self.lib.kmsc_o_merge_count(_ptr(a, u64p), len(a), _ptr(b, u64p), len(b))

    def pair_counts(self, offs_list, keys_list, key_bytes, n_buckets, bucket_ids=None, n_threads=1):
        n = len(offs_list)
        offs_list = [np.ascontiguousarray(o, np.int64) for o in offs_list]
        keys_list = [np.ascontiguousarray(k) for k in keys_list]
        op = (i64p * n)(*[_ptr(o, i64p) for o in offs_list])
        kp = (C.c_void_p * n)(*[k.ctypes.data for k in keys_list])
        out = np.zeros((n, n), np.int64)
        visits = C.c_int64()
        if bucket_ids is None:
            idp, nid = None, 0
        else:
            ids = np.ascontiguousarray(bucket_ids, np.int32)
            idp, nid = _ptr(ids, i32p), len(ids)
        self.lib.kmsc_o_pair_counts(op, kp, n, key_bytes, idp, nid, n_buckets, n_threads,
                                    _ptr(out, i64p), C.byref(visits))
        return out, visits.value

    # -- greedy arithmetic / DSU / streamvbyte ---------------------------------------
    def greedy_interval(self, n0): return self.lib.kmsc_o_greedy_interval(n0)
    def greedy_threshold(self, n0): return self.lib.kmsc_o_greedy_threshold(n0)
    def greedy_should_stop(self, total, updated, n0): return bool(self.lib.kmsc_o_greedy_should_stop(total, updated, n0))

    def greedy_argmax(self, w):
        w = np.ascontiguousarray(w, np.int64)
        j, k = C.c_int32(), C.c_int32()
        v = self.lib.kmsc_o_greedy_argmax(_ptr(w, i64p), w.shape[0], C.byref(j), C.byref(k))
        return v, j.value, k.value

    def bucket_histogram(self, codes, K, N, canonical=True):
        codes = np.ascontiguousarray(codes, np.uint8)
        hist = np.zeros(1 << N, np.int64)
        self.lib.kmsc_o_bucket_histogram(_ptr(codes, u8p), len(codes), K, N, int(canonical), _ptr(hist, i64p))
        return hist

    def mst(self, w):
        """(edges [(parent, child)], distances) of the `mst` driver over an exact intersection matrix"""
        w = np.ascontiguousarray(w, np.int64)
        n = w.shape[0]
        edges = np.zeros((max(1, n - 1), 2), np.int32)
        dist = np.zeros(max(1, n - 1), np.int64)
        ne = self.lib.kmsc_o_mst(_ptr(w, i64p), n, _ptr(edges, i32p), _ptr(dist, i64p))
        return edges[:ne].copy(), dist[:ne].copy()

    def svb_encode(self, vals):
        vals = np.ascontiguousarray(vals, np.uint32)
        out = np.zeros(self.lib.kmsc_o_svb0124_max_bytes(len(vals)) + 8, np.uint8)
        n = self.lib.kmsc_o_svb0124_encode(_ptr(vals, u32p), len(vals), _ptr(out, u8p))
        return out[:n].copy()

    # P6 container (no reference counterpart; layout in csrc/codec.cu)
    def codec_encode(self, K, N, key_bytes, offs, keys) -> bytes:
        offs = np.ascontiguousarray(offs, np.int64)
        keys = np.ascontiguousarray(keys, {2: np.uint16, 4: np.uint32, 8: np.uint64}[key_bytes])
        n = C.c_int64()
        p = self.lib.kmsc_o_codec_encode(K, N, key_bytes, offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                         keys.ctypes.data if len(keys) else None, C.byref(n))
        data = C.string_at(p, n.value)
        self.lib.kmsc_o_free(p)
        return data

    def codec_decode(self, data: bytes):
        buf = np.frombuffer(data, np.uint8)
        K, N, kb, nk = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        rc = self.lib.kmsc_o_codec_decode(_ptr(buf, u8p), len(buf), C.byref(K), C.byref(N), C.byref(kb), C.byref(nk), None, None)
        if rc != 0:
            raise ValueError("corrupt container")
        offs = np.zeros((1 << N.value) + 1, np.int64)
        keys = np.zeros(max(1, nk.value), {2: np.uint16, 4: np.uint32, 8: np.uint64}[kb.value])
        rc = self.lib.kmsc_o_codec_decode(_ptr(buf, u8p), len(buf), C.byref(K), C.byref(N), C.byref(kb), C.byref(nk),
                                          offs.ctypes.data_as(C.POINTER(C.c_int64)), keys.ctypes.data)
        if rc != 0:
            raise ValueError("corrupt container")
        return K.value, N.value, kb.value, offs, keys[:nk.value]

    def svb_decode(self, data, n):
        data = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(max(1, n), np.uint32)
        used = self.lib.kmsc_o_svb0124_decode(_ptr(data, u8p), _ptr(out, u32p), n)
        return out[:n].copy(), used


class Ref:
    """The reference's own code (oracle/_ref/libkmsc_ref.so). ``Ref.available()``
    is False where it was not prebuilt (it needs /root/reference to build)."""

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def __init__(self):
        L = self.lib = C.CDLL(str(REF_SO))
        L.ref_kmer_op.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_char_p, C.c_char, u64p, C.c_char_p]
        L.ref_bucket_key.argtypes = [C.c_int, C.c_uint64, i32p, u64p, u64p]
        L.ref_count_reads.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, C.c_int,
                                      u64p, u8p, i64p, u64p, i64p, i64p]
        L.ref_fasta.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, i64p]
        L.ref_sampled_set.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, i32p,
                                      C.c_int32, i64p, u64p, i64p, i64p]
        L.ref_set_from_spss.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, u64p,
                                        i64p, u64p]
        L.ref_spss_from_set.argtypes = [C.c_int, u64p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_void_p), i64p, i64p]
        L.ref_set_op.argtypes = [C.c_int, C.c_int, u64p, C.c_int64, u64p, C.c_int64, C.c_int, u64p, i64p, u64p]
        L.ref_kmer_set_set.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int32, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_double), C.POINTER(C.c_void_p), i64p, u64p, i32p,
                                       C.c_char_p]
        L.ref_split_stage.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_double), i64p]
        L.ref_reader_get.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int32, i64p, u64p, i32p]
        L.ref_random_ints.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, i32p]
        L.ref_seed_counter.restype = C.c_uint64
        L.ref_set_seed_counter.argtypes = [C.c_uint64]
        L.ref_dsu_new.argtypes = [C.c_int]
        L.ref_dsu_new.restype = C.c_void_p
        L.ref_dsu_free.argtypes = [C.c_void_p]
        L.ref_dsu_find.argtypes = [C.c_void_p, C.c_int]
        L.ref_dsu_same.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_dsu_unite.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_set_log_level.argtypes = [C.c_int]

    def kmer_op(self, cfg, op, bits=0, s=None, c="A"):
        K = CONFIGS[cfg][0]
        out = C.c_uint64()
        buf = C.create_string_buffer(K + 1)
        rc = self.lib.ref_kmer_op(cfg, op, bits, s.encode() if s is not None else None, c.encode(),
                                  C.byref(out), buf)
        assert rc == 0
        return out.value, buf.value.decode()

    def bucket_key(self, cfg, bits):
        b, k, back = C.c_int32(), C.c_uint64(), C.c_uint64()
        assert self.lib.ref_bucket_key(cfg, bits, C.byref(b), C.byref(k), C.byref(back)) == 0
        return b.value, k.value, back.value

    def add_with_max_u8(self, x, y): return self.lib.ref_add_with_max_u8(x, y)

    def count_reads(self, cfg, reads, canonical, cutoff, n_workers=1):
        K = CONFIGS[cfg][0]
        cap = max(1, sum(max(0, len(r) - K + 1) for r in reads))
        kmers = np.zeros(cap, np.uint64)
        counts = np.zeros(cap, np.uint8)
        kept = np.zeros(cap, np.uint64)
        nd, nk, cut = C.c_int64(), C.c_int64(), C.c_int64()
        rc = self.lib.ref_count_reads(cfg, _strs(reads), len(reads), int(canonical), n_workers, cutoff,
                                      _ptr(kmers, u64p), _ptr(counts, u8p), C.byref(nd), _ptr(kept, u64p),
                                      C.byref(nk), C.byref(cut))
        assert rc == 0, rc
        return kmers[:nd.value].copy(), counts[:nd.value].copy(), kept[:nk.value].copy(), cut.value

    def fasta(self, cfg, lines, canonical, n_workers=1):
        nd = C.c_int64()
        rc = self.lib.ref_fasta(cfg, _strs(lines), len(lines), int(canonical), n_workers, C.byref(nd))
        return rc, nd.value

    def sampled_set(self, cfg, strings, canonical, bucket_ids, n_workers=1):
        K = CONFIGS[cfg][0]
        ids = np.ascontiguousarray(bucket_ids, np.int32)
        offs = np.zeros(len(ids) + 1, np.int64)
        keys = np.zeros(max(1, sum(max(0, len(s) - K + 1) for s in strings)), np.uint64)
        size, weight = C.c_int64(), C.c_int64()
        rc = self.lib.ref_sampled_set(cfg, _strs(strings), len(strings), int(canonical), n_workers,
                                      _ptr(ids, i32p), len(ids), _ptr(offs, i64p), _ptr(keys, u64p),
                                      C.byref(size), C.byref(weight))
        assert rc == 0, rc
        return offs, keys[:offs[-1]].copy(), size.value, weight.value

    def set_from_spss(self, cfg, strings, canonical, n_workers=1):
        K = CONFIGS[cfg][0]
        out = np.zeros(max(1, sum(max(0, len(s) - K + 1) for s in strings)), np.uint64)
        n, h = C.c_int64(), C.c_uint64()
        rc = self.lib.ref_set_from_spss(cfg, _strs(strings), len(strings), int(canonical), n_workers,
                                        _ptr(out, u64p), C.byref(n), C.byref(h))
        assert rc == 0, rc
        return out[:n.value].copy(), h.value

    def spss_from_set(self, cfg, kmers, canonical, fast=True, n_workers=1):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        text = C.c_void_p()
        ns, w = C.c_int64(), C.c_int64()
        rc = self.lib.ref_spss_from_set(cfg, _ptr(kmers, u64p), len(kmers), int(canonical), int(fast),
                                        n_workers, C.byref(text), C.byref(ns), C.byref(w))
        assert rc == 0, rc
        s = C.string_at(text).decode()
        self.lib.ref_free(text)
        return [l for l in s.split("\n") if l != ""] if ns.value else [], w.value

    def set_op(self, cfg, op, a, b, n_workers=1):
        a = np.ascontiguousarray(a, np.uint64)
        b = np.ascontiguousarray(b, np.uint64)
        out = np.zeros(max(1, len(a) + len(b)), np.uint64)
        n, sc = C.c_int64(), C.c_uint64()
        code = {"add": 0, "sub": 1, "intersection": 2, "diff": 3, "hash": 4, "equals": 5}[op]
        rc = self.lib.ref_set_op(cfg, code, _ptr(a, u64p), len(a), _ptr(b, u64p), len(b), n_workers,
                                 _ptr(out, u64p), C.byref(n), C.byref(sc))
        assert rc == 0, rc
        return out[:n.value].copy() if code < 3 else sc.value

    def kmer_set_set(self, cfg, files, canonical, n_workers=1, stop_after_weights=False, dump_dir=""):
        n = len(files)
        phase = (C.c_double * 2)()
        log = C.c_void_p()
        sizes = np.zeros(max(1, n), np.int64)
        hashes = np.zeros(max(1, n), np.uint64)
        nn = C.c_int32()
        rc = self.lib.ref_kmer_set_set(cfg, _strs(files), n, int(canonical), n_workers,
                                       int(stop_after_weights), phase, C.byref(log), _ptr(sizes, i64p),
                                       _ptr(hashes, u64p), C.byref(nn), dump_dir.encode())
        text = C.string_at(log).decode() if log else ""
        if log:
            self.lib.ref_free(log)
        return dict(rc=rc, phase_s=(phase[0], phase[1]), log=text, sizes=sizes[:n].copy(),
                    hashes=hashes[:n].copy(), n_nodes=nn.value)

    def split_stage(self, cfg, file_j, file_k, canonical=True, n_workers=1):
        """one greedy-iteration split as the reference does it (kmer_set_set.h:332-343): seconds of
        (ToKmerSet x2, Intersection, Sub x2) and |j|, |k|, |n|, |j \\ n|, |k \\ n|"""
        sec = (C.c_double * 3)()
        sizes = np.zeros(5, np.int64)
        rc = self.lib.ref_split_stage(cfg, str(file_j).encode(), str(file_k).encode(), int(canonical), n_workers, sec, _ptr(sizes, i64p))
        assert rc == 0, rc
        return list(sec), sizes

    def reader_get(self, cfg, directory, canonical, i, n_workers=1):
        size, h, ns = C.c_int64(), C.c_uint64(), C.c_int32()
        rc = self.lib.ref_reader_get(cfg, str(directory).encode(), int(canonical), n_workers, i,
                                     C.byref(size), C.byref(h), C.byref(ns))
        assert rc == 0, rc
        return size.value, h.value, ns.value

    def random_ints(self, counter, n, lo, hi):
        out = np.zeros(max(1, n), np.int32)
        m = self.lib.ref_random_ints(counter, n, lo, hi, _ptr(out, i32p))
        return out[:m].copy()

    def seed_counter(self): return self.lib.ref_seed_counter()
    def set_seed_counter(self, c): self.lib.ref_set_seed_counter(c)


def set_ref_seed(seed: int) -> None:
    os.environ["KMSC_REF_SEED"] = str(seed)
