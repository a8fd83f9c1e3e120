"""ctypes binding of libkmsc (include/kmsc.h) -- the thin Python face of the C ABI.

Used by tests/ and bench.py. It holds no algorithm: every call goes straight into
the CUDA library and raises if the library or a GPU is missing (no CPU fallback).
Class and method names follow the reference's (KmerSet, KmerSetCompact.GetSampledKmerSet,
KmerSetSet's GetEdgeWeight ...), see include/kmsc.h for the file:line each replaces.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

import os

PKG_DIR = Path(__file__).resolve().parent
# KMSC_LIB selects an experimental build of the same library (tools/build_variants.sh)
LIB_PATH = Path(os.environ["KMSC_LIB"]) if os.environ.get("KMSC_LIB") else PKG_DIR / "libkmsc.so"

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)

KEY_DTYPES = {2: np.uint16, 4: np.uint32, 8: np.uint64}

EXPORTS = [
    "kmsc_last_error", "kmsc_version", "kmsc_ctx_create", "kmsc_ctx_destroy", "kmsc_ctx_sync",
    "kmsc_ctx_stream", "kmsc_ctx_launch_count", "kmsc_set_from_csr", "kmsc_set_from_kmers",
    "kmsc_set_to_csr", "kmsc_set_free", "kmsc_set_size", "kmsc_set_hash", "kmsc_set_info",
    "kmsc_set_from_spss", "kmsc_set_from_packed", "kmsc_sets_from_packed_batch", "kmsc_set_neighbors", "kmsc_spss_build", "kmsc_spss_fetch", "kmsc_spss_fetch_packed", "kmsc_comm_unique_id", "kmsc_comm_init", "kmsc_comm_destroy", "kmsc_comm_info", "kmsc_sets_exchange", "kmsc_set_bucket_offsets", "kmsc_set_export_range", "kmsc_set_import_range", "kmsc_pair_counts_stats", "kmsc_pair_counts_build", "kmsc_pair_counts", "kmsc_pair_counts_device", "kmsc_pair_counts_partial", "kmsc_pair_counts_rows",
    "kmsc_pair_split", "kmsc_pair_split_batch", "kmsc_set_union", "kmsc_set_diff", "kmsc_count_fasta", "kmsc_count_reads",
    "kmsc_count_get", "kmsc_count_last_counts", "kmsc_counter_create", "kmsc_counter_add_fasta",
    "kmsc_counter_add_reads", "kmsc_counter_prefetch", "kmsc_counter_finish", "kmsc_counter_free", "kmsc_bitmap_gram", "kmsc_codec_encode", "kmsc_codec_decode", "kmsc_free_host", "kmsc_host_alloc_pinned", "kmsc_host_free_pinned",
]


class KmscError(RuntimeError):
    pass


def build(verbose: bool = False) -> Path:
    """Compile libkmsc.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-s", "-j8", "-C", str(PKG_DIR / "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise KmscError("building libkmsc.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise KmscError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    # one NCCL per process: if this interpreter has (or will import) torch, the communicator must use
    # torch's bundled libnccl.so.2, not the system's (same SONAME, other version)
    if "KMSC_NCCL_LIB" not in os.environ:
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for base in (spec.submodule_search_locations if spec else []):
                cand = Path(base) / "lib" / "libnccl.so.2"
                if cand.exists():
                    os.environ["KMSC_NCCL_LIB"] = str(cand)
                    break
        except Exception:
            pass
    L = C.CDLL(str(LIB_PATH))
    L.kmsc_last_error.restype = C.c_char_p
    L.kmsc_version.restype = C.c_char_p
    L.kmsc_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.kmsc_ctx_destroy.argtypes = [C.c_void_p]
    L.kmsc_ctx_destroy.restype = None
    L.kmsc_ctx_sync.argtypes = [C.c_void_p]
    L.kmsc_ctx_stream.argtypes = [C.c_void_p]
    L.kmsc_ctx_stream.restype = C.c_void_p
    L.kmsc_ctx_launch_count.argtypes = [C.c_void_p]
    L.kmsc_ctx_launch_count.restype = C.c_int64
    L.kmsc_set_from_csr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _i64p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.kmsc_set_from_kmers.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _u64p, C.c_int64, C.POINTER(C.c_void_p)]
    L.kmsc_set_to_csr.argtypes = [C.c_void_p, C.c_void_p, _i64p, C.c_void_p]
    L.kmsc_set_free.argtypes = [C.c_void_p, C.c_void_p]
    L.kmsc_set_free.restype = None
    L.kmsc_set_size.argtypes = [C.c_void_p, C.c_void_p, _i64p]
    L.kmsc_set_hash.argtypes = [C.c_void_p, C.c_void_p, _u64p]
    L.kmsc_set_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _i64p]
    L.kmsc_set_from_spss.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, _i64p, C.c_int64,
                                     C.c_int, C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.kmsc_set_from_packed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, _i64p, C.c_int64,
                                       C.c_int, C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.kmsc_sets_from_packed_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32, C.POINTER(C.c_void_p),
                                              C.POINTER(C.c_void_p), _i64p, C.c_int, C.c_int, C.c_int32, C.c_int32,
                                              C.POINTER(C.c_void_p)]
    L.kmsc_pair_counts_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.kmsc_pair_counts.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, _i32p, C.c_int32, _i64p, _i64p]
    L.kmsc_pair_counts_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, _i32p, C.c_int32, C.c_void_p]
    L.kmsc_pair_counts_partial.argtypes = L.kmsc_pair_counts_device.argtypes
    L.kmsc_pair_counts_rows.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, _i32p, C.c_int32, _i32p,
                                        C.c_int32, _i64p]
    L.kmsc_pair_counts_build.argtypes = [C.c_void_p]
    L.kmsc_pair_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.kmsc_pair_split_batch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, _i64p,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.kmsc_set_union.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]
    L.kmsc_set_diff.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _i64p]
    L.kmsc_count_fasta.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p), _i64p, _i64p]
    L.kmsc_count_reads.argtypes = L.kmsc_count_fasta.argtypes
    L.kmsc_count_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    L.kmsc_count_last_counts.argtypes = [C.c_void_p, _u8p, C.c_int64]
    L.kmsc_set_neighbors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
    L.kmsc_spss_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.kmsc_spss_fetch.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
    L.kmsc_spss_fetch_packed.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    L.kmsc_set_bucket_offsets.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, _i64p]
    L.kmsc_set_export_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    L.kmsc_set_import_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.POINTER(C.c_void_p)]
    L.kmsc_comm_unique_id.argtypes = [C.c_void_p]
    L.kmsc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.kmsc_comm_destroy.argtypes = [C.c_void_p]
    L.kmsc_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.kmsc_sets_exchange.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, _i32p, C.POINTER(C.c_void_p), C.c_int32]
    L.kmsc_counter_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.kmsc_counter_add_fasta.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.kmsc_counter_add_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.kmsc_counter_prefetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.kmsc_counter_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _i64p, _i64p]
    L.kmsc_counter_free.argtypes = [C.c_void_p, C.c_void_p]
    L.kmsc_counter_free.restype = None
    L.kmsc_bitmap_gram.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, _i64p]
    L.kmsc_codec_encode.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), _i64p]
    L.kmsc_codec_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
    L.kmsc_free_host.argtypes = [C.c_void_p]
    L.kmsc_free_host.restype = None
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise KmscError(f"libkmsc error {rc}: {lib().kmsc_last_error().decode()}")


class DeviceSet:
    """Device-resident KmerSet<K,N,KeyType> in CSR form (owned handle)."""

    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.h = ctx, handle
        K, N, kb, n = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        _check(lib().kmsc_set_info(handle, C.byref(K), C.byref(N), C.byref(kb), C.byref(n)))
        self.K, self.N, self.key_bytes, self.n_keys = K.value, N.value, kb.value, n.value

    def Size(self) -> int:  # KmerSet::Size
        v = C.c_int64()
        _check(lib().kmsc_set_size(self.ctx.h, self.h, C.byref(v)))
        return v.value

    def Hash(self) -> int:  # KmerSet::Hash
        v = C.c_uint64()
        _check(lib().kmsc_set_hash(self.ctx.h, self.h, C.byref(v)))
        return v.value

    def to_csr(self):
        offs = np.zeros((1 << self.N) + 1, np.int64)
        keys = np.zeros(max(1, self.n_keys), KEY_DTYPES[self.key_bytes])
        _check(lib().kmsc_set_to_csr(self.ctx.h, self.h, offs.ctypes.data_as(_i64p), keys.ctypes.data))
        return offs, keys[: self.n_keys]

    def to_kmers(self) -> np.ndarray:
        """ascending 2K-bit k-mer values (what KmerSet::Find returns, sorted)."""
        offs, keys = self.to_csr()
        buckets = np.repeat(np.arange(1 << self.N, dtype=np.uint64), np.diff(offs))
        return (buckets << np.uint64(2 * self.K - self.N)) | keys.astype(np.uint64)

    def free(self):
        if self.h:
            lib().kmsc_set_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU + one stream (kmsc_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        _check(lib().kmsc_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h.value
        self.device = device

    def close(self):
        if self.h:
            lib().kmsc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self): _check(lib().kmsc_ctx_sync(self.h))
    def launch_count(self) -> int: return lib().kmsc_ctx_launch_count(self.h)
    def stream(self) -> int: return lib().kmsc_ctx_stream(self.h) or 0

    # -- sets -----------------------------------------------------------------
    def set_from_csr(self, K, N, key_bytes, offs, keys) -> DeviceSet:
        offs = np.ascontiguousarray(offs, np.int64)
        keys = np.ascontiguousarray(keys, KEY_DTYPES[key_bytes])
        h = C.c_void_p()
        _check(lib().kmsc_set_from_csr(self.h, K, N, key_bytes, offs.ctypes.data_as(_i64p), keys.ctypes.data, C.byref(h)))
        return DeviceSet(self, h.value)

    def set_from_kmers(self, K, N, key_bytes, kmers) -> DeviceSet:
        kmers = np.ascontiguousarray(kmers, np.uint64)
        h = C.c_void_p()
        _check(lib().kmsc_set_from_kmers(self.h, K, N, key_bytes, kmers.ctypes.data_as(_u64p), len(kmers), C.byref(h)))
        return DeviceSet(self, h.value)

    def set_from_spss(self, K, N, key_bytes, strings, canonical=True, dedup=True, bucket_lo=0, bucket_hi=None,
                      text=None, str_offs=None) -> DeviceSet:
        """KmerSetCompact::ToKmerSet (dedup) / GetSampledKmerSet over a bucket range (dedup=False)."""
        if text is None:
            bs = [s.encode() if isinstance(s, str) else s for s in strings]
            str_offs = np.zeros(len(bs) + 1, np.int64)
            np.cumsum([len(b) for b in bs], out=str_offs[1:])
            text = np.frombuffer(b"".join(bs), np.uint8) if bs else np.zeros(0, np.uint8)
        text = np.ascontiguousarray(text, np.uint8)
        str_offs = np.ascontiguousarray(str_offs, np.int64)
        h = C.c_void_p()
        hi = (1 << N) if bucket_hi is None else bucket_hi
        _check(lib().kmsc_set_from_spss(self.h, K, N, key_bytes, text.ctypes.data, str_offs.ctypes.data_as(_i64p),
                                        len(str_offs) - 1, int(canonical), int(dedup), bucket_lo, hi, C.byref(h)))
        return DeviceSet(self, h.value)

    def set_from_packed(self, K, N, key_bytes, words, str_offs, canonical=True, dedup=True, bucket_lo=0,
                        bucket_hi=None, words_ptr=None) -> DeviceSet:
        """same decode from 2-bit packed bases (32 per uint64, first base in the top bits);
        words_ptr: raw host pointer (e.g. a pinned torch tensor's data_ptr) instead of an array"""
        str_offs = np.ascontiguousarray(str_offs, np.int64)
        if words_ptr is None:
            words = np.ascontiguousarray(words, np.uint64)
            words_ptr = words.ctypes.data
        h = C.c_void_p()
        hi = (1 << N) if bucket_hi is None else bucket_hi
        _check(lib().kmsc_set_from_packed(self.h, K, N, key_bytes, C.c_void_p(words_ptr), str_offs.ctypes.data_as(_i64p),
                                          len(str_offs) - 1, int(canonical), int(dedup), bucket_lo, hi, C.byref(h)))
        return DeviceSet(self, h.value)

    def sets_from_packed_batch(self, K, N, key_bytes, words_list, str_offs_list, canonical=True, dedup=True,
                               bucket_lo=0, bucket_hi=None, words_ptrs=None):
        """m packed SPSS -> m device sets in one batched launch sequence (KmerSetSet ctor's per-set loop);
        words_ptrs: raw host pointers (e.g. pinned torch tensors) instead of arrays"""
        m = len(str_offs_list)
        offs = [np.ascontiguousarray(o, np.int64) for o in str_offs_list]
        if words_ptrs is None:
            words_list = [np.ascontiguousarray(w, np.uint64) for w in words_list]
            words_ptrs = [w.ctypes.data for w in words_list]
        wp = (C.c_void_p * m)(*words_ptrs)
        op = (C.c_void_p * m)(*[o.ctypes.data for o in offs])
        ns = np.array([len(o) - 1 for o in offs], np.int64)
        out = (C.c_void_p * m)()
        hi = (1 << N) if bucket_hi is None else bucket_hi
        _check(lib().kmsc_sets_from_packed_batch(self.h, K, N, key_bytes, m, wp, op, ns.ctypes.data_as(_i64p), int(canonical),
                                                 int(dedup), bucket_lo, hi, out))
        return [DeviceSet(self, out[j]) for j in range(m)]

    def pair_counts_stats(self) -> dict:
        out = (C.c_double * 8)()
        _check(lib().kmsc_pair_counts_stats(self.h, out))
        keys = ["main_ms", "plan_ms", "keys", "distinct", "retries", "L", "main_launches", "algo_bytes"]
        return dict(zip(keys, list(out)))

    # -- P3 -----------------------------------------------------------------------
    def _handles(self, sets):
        return (C.c_void_p * len(sets))(*[s.h for s in sets])

    def pair_counts(self, sets, bucket_ids=None, with_visits=False):
        n = len(sets)
        out = np.zeros((n, n), np.int64)
        visits = C.c_int64()
        if bucket_ids is None:
            idp, nid = None, 0
        else:
            ids = np.ascontiguousarray(bucket_ids, np.int32)
            idp, nid = ids.ctypes.data_as(_i32p), len(ids)
        _check(lib().kmsc_pair_counts(self.h, self._handles(sets), n, idp, nid, out.ctypes.data_as(_i64p), C.byref(visits)))
        return (out, visits.value) if with_visits else out

    def pair_counts_device(self, sets, d_out_ptr: int, bucket_ids=None, partial=False):
        """matrix left on the device; partial=True: this rank's own counts, no all-reduce"""
        n = len(sets)
        if bucket_ids is None:
            idp, nid = None, 0
        else:
            ids = np.ascontiguousarray(bucket_ids, np.int32)
            idp, nid = ids.ctypes.data_as(_i32p), len(ids)
        fn = lib().kmsc_pair_counts_partial if partial else lib().kmsc_pair_counts_device
        _check(fn(self.h, self._handles(sets), n, idp, nid, C.c_void_p(d_out_ptr)))

    def pair_counts_build(self) -> int:
        """0 = hash build, 1 = merge build (main phase of the last pair_counts call)"""
        return int(lib().kmsc_pair_counts_build(self.h))

    def pair_counts_rows(self, sets, rows, bucket_ids=None):
        n = len(sets)
        rows = np.ascontiguousarray(rows, np.int32)
        out = np.zeros((len(rows), n), np.int64)
        if bucket_ids is None:
            idp, nid = None, 0
        else:
            ids = np.ascontiguousarray(bucket_ids, np.int32)
            idp, nid = ids.ctypes.data_as(_i32p), len(ids)
        _check(lib().kmsc_pair_counts_rows(self.h, self._handles(sets), n, rows.ctypes.data_as(_i32p), len(rows),
                                           idp, nid, out.ctypes.data_as(_i64p)))
        return out

    # -- P4 -------------------------------------------------------------------------
    def set_neighbors(self, s: DeviceSet, canonical=True) -> np.ndarray:
        out = np.zeros((max(1, s.n_keys), 8), np.int32)
        _check(lib().kmsc_set_neighbors(self.h, s.h, int(canonical), out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out[: s.n_keys]

    def spss_build(self, s: DeviceSet, canonical=True, rounds=0, fetch=True):
        """SPSS of a set, built on the device: list of strings (fetch=False: just (n_strings, n_chars))"""
        ns, nc = C.c_int64(0), C.c_int64(0)
        _check(lib().kmsc_spss_build(self.h, s.h, int(canonical), int(rounds), C.byref(ns), C.byref(nc)))
        if not fetch:
            return ns.value, nc.value
        text = C.create_string_buffer(max(1, nc.value))
        offs = np.zeros(ns.value + 1, np.int64)
        _check(lib().kmsc_spss_fetch(self.h, text, offs.ctypes.data_as(_i64p)))
        raw = text.raw
        return [raw[offs[i]:offs[i + 1]].decode() for i in range(ns.value)]

    def spss_build_packed(self, s: DeviceSet, canonical=True, rounds=0):
        """SPSS of a set as (words, str_offs): the packed container kmsc_set_from_packed reads"""
        ns, nc = C.c_int64(0), C.c_int64(0)
        _check(lib().kmsc_spss_build(self.h, s.h, int(canonical), int(rounds), C.byref(ns), C.byref(nc)))
        words = np.zeros((nc.value + 31) // 32 + 2, np.uint64)
        offs = np.zeros(ns.value + 1, np.int64)
        _check(lib().kmsc_spss_fetch_packed(self.h, words.ctypes.data_as(_u64p), offs.ctypes.data_as(_i64p)))
        return words, offs

    # -- multi-GPU inside the library -----------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().kmsc_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, n_ranks: int, id128: bytes):
        assert len(id128) == 128
        _check(lib().kmsc_comm_init(self.h, rank, n_ranks, C.c_char_p(id128)))

    def comm_destroy(self):
        _check(lib().kmsc_comm_destroy(self.h))

    def comm_info(self):
        r, n = C.c_int(), C.c_int()
        _check(lib().kmsc_comm_info(self.h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def sets_exchange(self, mine, cuts, n_total):
        """whole sets decoded by this rank -> all n_total sets restricted to this rank's bucket range"""
        cuts = np.ascontiguousarray(cuts, np.int32)
        out = (C.c_void_p * n_total)()
        _check(lib().kmsc_sets_exchange(self.h, self._handles(mine) if mine else None, len(mine), cuts.ctypes.data_as(_i32p), out, n_total))
        return [DeviceSet(self, out[g]) for g in range(n_total)]

    # -- multi-GPU exchange helpers ---------------------------------------------------------
    def set_bucket_offsets(self, s: DeviceSet, buckets) -> np.ndarray:
        b = np.ascontiguousarray(buckets, np.int32)
        out = np.zeros(len(b), np.int64)
        _check(lib().kmsc_set_bucket_offsets(self.h, s.h, b.ctypes.data_as(C.POINTER(C.c_int32)), len(b), out.ctypes.data_as(_i64p)))
        return out

    def set_export_range(self, s: DeviceSet, bucket_lo, bucket_hi, key_lo, key_hi, d_offs_ptr, d_keys_ptr):
        _check(lib().kmsc_set_export_range(self.h, s.h, bucket_lo, bucket_hi, key_lo, key_hi, d_offs_ptr, d_keys_ptr))

    def set_import_range(self, K, N, key_bytes, bucket_lo, bucket_hi, d_offs_ptr, d_keys_ptr, n_keys) -> DeviceSet:
        h = C.c_void_p()
        _check(lib().kmsc_set_import_range(self.h, K, N, key_bytes, bucket_lo, bucket_hi, d_offs_ptr, d_keys_ptr, n_keys, C.byref(h)))
        return DeviceSet(self, h.value)

    def pair_split(self, j: DeviceSet, k: DeviceSet, want_inter: bool = True):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().kmsc_pair_split(self.h, j.h, k.h, C.byref(a) if want_inter else None, C.byref(b), C.byref(c)))
        return (DeviceSet(self, a.value) if want_inter else None), DeviceSet(self, b.value), DeviceSet(self, c.value)

    def pair_split_batch(self, js, ks, inter_hint=None, want_inter=True, want_j=True, want_k=True):
        """m pairs in one streaming pass -> (inter, j_minus, k_minus), each a list of m sets or None."""
        m = len(js)
        assert len(ks) == m
        arrs = [(C.c_void_p * m)() if w else None for w in (want_inter, want_j, want_k)]
        hint = None
        if inter_hint is not None:
            hint = np.ascontiguousarray(inter_hint, dtype=np.int64)
            assert hint.shape == (m,)
        _check(lib().kmsc_pair_split_batch(self.h, self._handles(js), self._handles(ks), m,
                                           hint.ctypes.data_as(_i64p) if hint is not None else None, *arrs))
        return tuple([DeviceSet(self, a[i]) for i in range(m)] if a is not None else None for a in arrs)

    def set_union(self, sets) -> DeviceSet:
        h = C.c_void_p()
        _check(lib().kmsc_set_union(self.h, self._handles(sets), len(sets), C.byref(h)))
        return DeviceSet(self, h.value)

    def set_diff(self, a: DeviceSet, b: DeviceSet) -> int:
        v = C.c_int64()
        _check(lib().kmsc_set_diff(self.h, a.h, b.h, C.byref(v)))
        return v.value

    # -- P1 ---------------------------------------------------------------------------
    def _count(self, fn, K, N, key_bytes, data: bytes, canonical, cutoff):
        buf = np.frombuffer(data, np.uint8)
        h = C.c_void_p()
        cut, nd = C.c_int64(), C.c_int64()
        _check(fn(self.h, K, N, key_bytes, buf.ctypes.data, len(buf), int(canonical), int(cutoff), C.byref(h),
                  C.byref(cut), C.byref(nd)))
        return DeviceSet(self, h.value), cut.value, nd.value

    def count_fasta(self, K, N, key_bytes, data: bytes, canonical=True, cutoff=1):
        return self._count(lib().kmsc_count_fasta, K, N, key_bytes, data, canonical, cutoff)

    def count_reads(self, K, N, key_bytes, data: bytes, canonical=True, cutoff=1):
        return self._count(lib().kmsc_count_reads, K, N, key_bytes, data, canonical, cutoff)

    def count_get(self, kmer: int) -> int:
        v = C.c_int()
        _check(lib().kmsc_count_get(self.h, kmer, C.byref(v)))
        return v.value

    def count_chunks(self, K, N, key_bytes, chunks, canonical=True, cutoff=1, fasta=True, prefetch=True):
        """streaming counter: every chunk holds whole records; returns (set, cutoff_count, n_distinct)"""
        c = C.c_void_p()
        _check(lib().kmsc_counter_create(self.h, K, N, key_bytes, int(canonical), C.byref(c)))
        try:
            add = lib().kmsc_counter_add_fasta if fasta else lib().kmsc_counter_add_reads
            # bytes, or uint8 numpy views (e.g. of pinned memory); the copy of chunk i + 1 is announced before
            # chunk i is counted (kmsc_counter_prefetch)
            ptrs = [ch.ctypes.data if isinstance(ch, np.ndarray) else C.cast(C.c_char_p(ch), C.c_void_p).value for ch in chunks]
            for i, ch in enumerate(chunks):
                if prefetch and i + 1 < len(chunks) and not os.environ.get("KMSC_NO_PREFETCH"):
                    _check(lib().kmsc_counter_prefetch(self.h, c, C.c_void_p(ptrs[i + 1]), len(chunks[i + 1])))
                _check(add(self.h, c, C.c_void_p(ptrs[i]), len(ch)))
            h, cut, nd = C.c_void_p(), C.c_int64(), C.c_int64()
            _check(lib().kmsc_counter_finish(self.h, c, cutoff, C.byref(h), C.byref(cut), C.byref(nd)))
        finally:
            lib().kmsc_counter_free(self.h, c)
        return DeviceSet(self, h.value), cut.value, nd.value

    def count_last_counts(self, n: int) -> np.ndarray:
        out = np.zeros(max(1, n), np.uint8)
        _check(lib().kmsc_count_last_counts(self.h, out.ctypes.data_as(_u8p), n))
        return out[:n]

    # -- P5 / P6 ------------------------------------------------------------------------
    def bitmap_gram(self, sets):
        n = len(sets)
        out = np.zeros((n, n), np.int64)
        _check(lib().kmsc_bitmap_gram(self.h, self._handles(sets), n, out.ctypes.data_as(_i64p)))
        return out

    def codec_encode(self, s: DeviceSet) -> bytes:
        p, n = C.c_void_p(), C.c_int64()
        _check(lib().kmsc_codec_encode(self.h, s.h, C.byref(p), C.byref(n)))
        data = C.string_at(p, n.value)
        lib().kmsc_free_host(p)
        return data

    def codec_decode(self, data: bytes) -> DeviceSet:
        buf = np.frombuffer(data, np.uint8)
        h = C.c_void_p()
        _check(lib().kmsc_codec_decode(self.h, buf.ctypes.data, len(buf), C.byref(h)))
        return DeviceSet(self, h.value)
