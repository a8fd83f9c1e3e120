/* kmsc.h -- C ABI of libkmsc, the B200 (sm_100a) data-parallel core for
 * kkty/kmer-sets-compression's `kmerset-multiple-compress` hot path.
 *
 * The reference is a header-only C++17 template library with no FFI of its own
 * (SURVEY.md section 8b); the drop-in boundary is therefore this C ABI, bound
 * by the C++17 facade in kmer-sets-compression_b200/host/ whose classes mirror
 * the reference's (KmerSet, KmerCounter, KmerSetCompact, KmerSetSet). Every
 * entry point names the reference interface it replaces (paths relative to
 * /root/reference/). Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative KMSC_E_* code otherwise;
 *    kmsc_last_error() returns the calling thread's last message.
 *  - There is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with KMSC_E_CUDA.
 *  - A k-mer is the reference's 2K-bit value (lib/core/kmer.h:22-46): A=0 C=1
 *    G=2 T=3, first base most significant. bucket = value >> (2K-N),
 *    key = value mod 2^(2K-N) (lib/core/kmer_set.h:22-43).
 *  - A device set (kmsc_set) is the CSR form of the reference's bucketed set:
 *    offs[2^N + 1] and keys[] ascending inside each bucket. key_bytes is
 *    sizeof(KeyType): 2, 4 or 8 (uint16/uint32/uint64 as in
 *    src/kmerset-multiple-compress.cc:149-157; uint64 for K=31).
 *  - Host buffers belong to the caller; device buffers belong to the library
 *    until the matching *_free. Calls that return host data are synchronous at
 *    return. Calls that only produce device sets (kmsc_sets_from_packed_batch,
 *    kmsc_sets_exchange, kmsc_pair_split*) may return with work still queued on the
 *    context's stream: the sets are valid for every later call on that context,
 *    and kmsc_ctx_sync waits for them. kmsc_pair_counts_device synchronises once
 *    internally (tile statistics) and leaves the matrix on the stream.
 *  - One kmsc_ctx = one GPU + one stream; drive it from one host thread.
 *    Multi-GPU = one context per process/rank (kmsc_comm_init), sets restricted
 *    to that rank's bucket range, partial matrices summed inside kmsc_pair_counts*.
 */
#ifndef KMSC_H_
#define KMSC_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMSC_OK 0
#define KMSC_E_INVALID (-1) /* bad argument */
#define KMSC_E_CUDA (-2)    /* CUDA runtime/driver error, or no device */
#define KMSC_E_NOMEM (-3)
#define KMSC_E_FORMAT (-4)  /* malformed input text (FASTA / SPSS) */
#define KMSC_E_STATE (-5)

typedef struct kmsc_ctx kmsc_ctx;
typedef struct kmsc_set kmsc_set;

const char* kmsc_last_error(void);
const char* kmsc_version(void);

/* ---- context ---------------------------------------------------------------- */
/* stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL to
 * let the library create its own non-blocking stream. */
int kmsc_ctx_create(int device, void* stream, kmsc_ctx** out);
void kmsc_ctx_destroy(kmsc_ctx* ctx);
int kmsc_ctx_sync(kmsc_ctx* ctx);
void* kmsc_ctx_stream(kmsc_ctx* ctx);
/* kernels launched by this context so far (bench.py's gpu_launches) */
int64_t kmsc_ctx_launch_count(kmsc_ctx* ctx);

/* ---- device sets: KmerSet<K,N,KeyType> as CSR --------------------------------- */
/* Host CSR -> device set. offs: int64[2^N+1]; keys: key_bytes each, ascending
 * within each bucket (duplicates allowed, as GetSampledKmerSet may produce,
 * lib/core/kmer_set_compact.h:120-203). Replaces building a KmerSet by Add()
 * (lib/core/kmer_set.h:77-83). */
int kmsc_set_from_csr(kmsc_ctx* ctx, int K, int N, int key_bytes, const int64_t* offs,
                      const void* keys, kmsc_set** out);
/* Same, from ascending 2K-bit k-mer values (uint64), e.g. KmerSet::Find output. */
int kmsc_set_from_kmers(kmsc_ctx* ctx, int K, int N, int key_bytes, const uint64_t* kmers,
                        int64_t n, kmsc_set** out);
/* device set -> host CSR (either pointer may be NULL). */
int kmsc_set_to_csr(kmsc_ctx* ctx, const kmsc_set* set, int64_t* offs, void* keys);
void kmsc_set_free(kmsc_ctx* ctx, kmsc_set* set);
/* KmerSet::Size (lib/core/kmer_set.h:62-68) and KmerSet::Hash = XOR of all 2K-bit
 * values (lib/core/kmer_set.h:224-244, kmer.h:211), computed on the device. */
int kmsc_set_size(kmsc_ctx* ctx, const kmsc_set* set, int64_t* size);
int kmsc_set_hash(kmsc_ctx* ctx, const kmsc_set* set, uint64_t* hash);
int kmsc_set_info(const kmsc_set* set, int* K, int* N, int* key_bytes, int64_t* n_keys);

/* ---- multi-GPU exchange: bucket ranges of a set ---------------------------------------- */
/* A rank that decoded WHOLE sets hands every rank the part of each set inside that rank's
 * bucket range (the sets are sorted by bucket, so a range is one contiguous key slice):
 * the decode work is split over the ranks by set, one all-to-all moves the slices, and
 * every rank imports all n sets restricted to its prefix range.
 * offsets: out[i] = index of the first key of bucket buckets[i] (buckets[i] in [0, 2^N]). */
int kmsc_set_bucket_offsets(kmsc_ctx* ctx, const kmsc_set* set, const int32_t* buckets, int32_t n, int64_t* out);
/* device -> device on the context's stream (no synchronisation): d_offs receives
 * bucket_hi - bucket_lo + 1 offsets rebased to 0, d_keys the keys [key_lo, key_hi) =
 * the offsets of bucket_lo / bucket_hi from kmsc_set_bucket_offsets. */
int kmsc_set_export_range(kmsc_ctx* ctx, const kmsc_set* set, int32_t bucket_lo, int32_t bucket_hi,
                          int64_t key_lo, int64_t key_hi, uint32_t* d_offs, void* d_keys);
/* the inverse: a set whose buckets outside [bucket_lo, bucket_hi) are empty; d_offs / d_keys
 * are device memory in the layout kmsc_set_export_range writes. */
int kmsc_set_import_range(kmsc_ctx* ctx, int K, int N, int key_bytes, int32_t bucket_lo, int32_t bucket_hi,
                          const uint32_t* d_offs, const void* d_keys, int64_t n_keys, kmsc_set** out);

/* ---- multi-GPU inside the library: one context per rank, NCCL communicator owned by the context ---- */
/* The reference is one shared-memory process (boost::asio::thread_pool, 40 call sites); what shards is
 * the sum over buckets of lib/core/kmer_set_set.h:161-181. With a communicator on the context,
 * kmsc_pair_counts / _device / _rows all-reduce the per-rank partial matrices INSIDE the call (one
 * ncclAllReduce of n*n int64 on the context's stream), so every rank returns the full matrix.
 * NCCL is loaded at run time (libnccl.so.2); without it these calls fail with KMSC_E_STATE.
 * id128: 128 bytes from kmsc_comm_unique_id on one rank, handed to every rank by the host program
 * (MPI, torch.distributed, a file, or plain memory between threads). */
int kmsc_comm_unique_id(void* id128);
int kmsc_comm_init(kmsc_ctx* ctx, int rank, int n_ranks, const void* id128);
int kmsc_comm_destroy(kmsc_ctx* ctx);
int kmsc_comm_info(kmsc_ctx* ctx, int* rank, int* n_ranks);
/* Every rank decoded n_mine WHOLE sets (global set index = rank + j * n_ranks for its j-th set) and
 * receives all n_total = n_mine * n_ranks sets restricted to ITS bucket range [cuts[rank],
 * cuts[rank+1]) (cuts: n_ranks + 1 ascending bucket indices from 0 to 2^N, the same on all ranks).
 * One table all-gather + one host synchronisation (the sizes), then one grouped ncclSend / ncclRecv
 * moves every slice straight between the sets' key arrays. Replaces kmsc_set_bucket_offsets /
 * _export_range / _import_range + a caller-side all-to-all. out: n_total handles. */
int kmsc_sets_exchange(kmsc_ctx* ctx, const kmsc_set* const* mine, int32_t n_mine, const int32_t* cuts,
                       kmsc_set** out, int32_t n_total);

/* ---- P2: SPSS text -> device set ------------------------------------------------ */
/* Replaces KmerSetCompact::GetSampledKmerSet (lib/core/kmer_set_compact.h:120-203;
 * dedup = 0: duplicates kept) and KmerSetCompact::ToKmerSet / GetKmerSetFromSPSS
 * (lib/core/spss.h:1861-1941; dedup = 1: hash-set semantics). text = the strings
 * concatenated (A/C/G/T only), str_offs: int64[n_strings + 1]. Only k-mers whose
 * bucket lies in [bucket_lo, bucket_hi) are kept (0, 2^N = all): a rank's shard. */
int kmsc_set_from_spss(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* text,
                       const int64_t* str_offs, int64_t n_strings, int canonical, int dedup,
                       int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out);

/* Same decode from the 2-bit packed form KmerSetCompact keeps in memory
 * (lib/core/kmer_set_compact.h:206-255, 338-347): words = 32 bases per uint64,
 * first base in the top two bits, codes A=0 C=1 G=2 T=3 (the bit order inside
 * std::vector<bool> is not observable, so the container is ours); str_offs in
 * BASES. words must hold ceil(total_bases / 32) + 1 entries. */
int kmsc_set_from_packed(kmsc_ctx* ctx, int K, int N, int key_bytes, const uint64_t* words,
                         const int64_t* str_offs, int64_t n_strings, int canonical, int dedup,
                         int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out);

/* The same decode for m sets in ONE batch -- the per-set loop of KmerSetSet's constructor
 * (lib/core/kmer_set_set.h:138-153: one GetSampledKmerSet task per set). Device work is
 * batched over the sets (count, partition, sort and offset kernels cover a group of sets per
 * launch), the host-to-device copies of later groups overlap the kernels of earlier ones,
 * and the host synchronises twice per group instead of three times per set. words[j] /
 * str_offs[j] / n_strings[j] as for kmsc_set_from_packed; pinned host memory makes the
 * copies asynchronous. out: m handles. On error no set is returned. */
int kmsc_sets_from_packed_batch(kmsc_ctx* ctx, int K, int N, int key_bytes, int32_t m,
                                const uint64_t* const* words, const int64_t* const* str_offs,
                                const int64_t* n_strings, int canonical, int dedup,
                                int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out);

/* SPSS construction support: the de Bruijn neighbours of every k-mer of a set.
 * out (host, 8 * n_keys int32): out[8 i + c] = Kmer::Next(base c) of k-mer i, out[8 i + 4 + c] =
 * Kmer::Prev(base c) (lib/core/kmer.h:136-186), each -1 if absent from the set, else
 * (index << 1) | flip, index = position in the set's key order (the order of kmsc_set_to_csr),
 * flip = 1 if the set holds the reverse complement (canonical != 0). Replaces the hash-set
 * Contains() calls of the reference's unitig / path-cover construction
 * (lib/core/spss.h:230-615, 1039-1858) by one device binary search per neighbour. */
int kmsc_set_neighbors(kmsc_ctx* ctx, const kmsc_set* set, int canonical, int32_t* out);

/* SPSS construction on the device. Replaces GetSPSS / GetSPSSCanonical
 * (lib/core/spss.h:230-615 unitigs, :1039-1858 greedy path cover) and with them
 * KmerSetCompact::FromKmerSet (lib/core/kmer_set_compact.h:89-100, lib/core/spss.h:1836-1858): a set of
 * strings that spells every k-mer of the set exactly once (canonical != 0: it or its reverse complement),
 * the property the reference's tests check (test/spss.cc:57-68, 113-124). Every k-mer port (left / right
 * end) is linked to at most one neighbouring port by `rounds` rounds of mutual proposals (<= 0: 8; round
 * one joins exactly the unitig ends, later rounds stitch unitigs like the reference's path cover), the
 * resulting paths are ranked by pointer jumping, cycles are cut at their smallest k-mer. Strings are
 * ordered by their first k-mer's position in the set; the output is deterministic for a given set.
 * kmsc_spss_build leaves the text on the device and reports its size; kmsc_spss_fetch copies the result
 * of the LAST build of this context: text (n_chars bytes, no separators) and str_offs (n_strings + 1). */
int kmsc_spss_build(kmsc_ctx* ctx, const kmsc_set* set, int canonical, int rounds, int64_t* n_strings,
                    int64_t* n_chars);
int kmsc_spss_fetch(kmsc_ctx* ctx, char* text, int64_t* str_offs);
/* The same result in the container KmerSetCompact holds (lib/core/kmer_set_compact.h:206-255): words
 * ((n_chars + 31) / 32 uint64: 2 bits per base, 32 bases per word, first base in the top bits, strings back
 * to back -- what kmsc_set_from_packed reads) and str_offs in bases. Packed on the device: the text never
 * crosses the bus. */
int kmsc_spss_fetch_packed(kmsc_ctx* ctx, uint64_t* words, int64_t* str_offs);

/* ---- P3: all-pairs intersection counts ------------------------------------------ */
/* Replaces GetEdgeWeight and the initial all-pairs loop of KmerSetSet's
 * constructor (lib/core/kmer_set_set.h:158-219): out[i*n + j] =
 * sum over b in bucket_ids of |merge-count(S_i[b], S_j[b])| (int64), for ALL i, j
 * (symmetric; the diagonal holds each set's key count over those buckets).
 * bucket_ids == NULL: all 2^N buckets (exact matrix); otherwise the reference's
 * sampled list (any order; an id listed twice counts once, as its map does,
 * :127-131). out is host memory (n*n int64). key_visits (may be NULL) receives
 * sum_{i<j} sum_b (len_i[b] + len_j[b]) -- the work the reference's merge does.
 * Any n: up to 256 sets are one pass; more are covered by pairs of 128-set groups. */
int kmsc_pair_counts(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                     const int32_t* bucket_ids, int32_t n_ids, int64_t* out, int64_t* key_visits);
/* Same, result left in DEVICE memory d_out (n*n int64, on the context's stream). On a context with
 * a communicator (kmsc_comm_init) the per-rank partial matrices are summed by one ncclAllReduce
 * inside the call: every rank ends with the full matrix. */
int kmsc_pair_counts_device(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                            const int32_t* bucket_ids, int32_t n_ids, int64_t* d_out);
/* This rank's PARTIAL matrix only (no all-reduce even with a communicator), in device memory:
 * |S_i & S_j| inside the rank's bucket range = the exact inter_hint of kmsc_pair_split_batch on
 * a prefix shard. */
int kmsc_pair_counts_partial(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                             const int32_t* bucket_ids, int32_t n_ids, int64_t* d_out);
/* Row mode (lib/core/kmer_set_set.h:385-425, the 3n-2 re-weights after a merge):
 * out[r*n + l] = weight(rows[r], l) for l in [0, n). Every column set is read once and merged
 * (the reference's two-pointer loop, :165-180: duplicates count with min multiplicity) against
 * the row sets' runs staged in shared memory -- 3 rows cost one pass over the sets, not a full
 * n x n matrix. All-reduced over the ranks like kmsc_pair_counts. */
int kmsc_pair_counts_rows(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                          const int32_t* rows, int32_t n_rows, const int32_t* bucket_ids,
                          int32_t n_ids, int64_t* out);

/* Device-side timing and tiling facts of the LAST kmsc_pair_counts* call, for
 * bench.py's roofline: out[0] main-kernel ms (CUDA events on the context's
 * stream, summed over the call's launches), out[1] planning-kernels ms, out[2]
 * keys read, out[3] distinct keys seen, out[4] tiles re-run after a table
 * overflow, out[5] tile target L, out[6] main-kernel launches, out[7] algorithmic
 * bytes (keys * sizeof(KeyType) + offsets read + n*n*8). */
int kmsc_pair_counts_stats(kmsc_ctx* ctx, double* out8);
/* Which build the main phase of the LAST kmsc_pair_counts* call used: 0 = shared-memory hash
 * table (cost per key), 1 = warp-wide multiway merge (cost per distinct key; chosen for up to
 * 128 related sets), 2 = warp-private hash tables, 3 = lane-private tables (2 and 3: up to 64 sets,
 * measured slower, never chosen by the library). Same results; the environment variable
 * KMSC_P3_BUILD=hash|merge|whash|lane forces one. */
int kmsc_pair_counts_build(kmsc_ctx* ctx);

/* ---- P4: pair split / set algebra ------------------------------------------------ */
/* Replaces kmer_set_set.h:332-343: n = Intersection(j, k); j.Sub(n); k.Sub(n)
 * (lib/core/kmer_set.h:177-187, 301-305) in one pass over both CSRs. Any output
 * pointer may be NULL. Inputs must be duplicate-free (true sets). */
int kmsc_pair_split(kmsc_ctx* ctx, const kmsc_set* j, const kmsc_set* k, kmsc_set** inter,
                    kmsc_set** j_minus, kmsc_set** k_minus);
/* The same for m pairs in ONE streaming pass (every input key read once, every output key
 * written once): the n-1 tree edges of the `mst` driver, or one greedy iteration. All sets share
 * (K, N, KeyType). inter / j_minus / k_minus: arrays of m handles, or NULL for an output that is
 * not wanted. inter_hint (may be NULL): inter_hint[p] = |js[p] & ks[p]| if the caller knows it
 * (kmsc_pair_counts over all buckets gives it); the outputs are then allocated exactly and
 * written directly. A wrong hint costs a second pass for that pair, never a wrong result. */
int kmsc_pair_split_batch(kmsc_ctx* ctx, const kmsc_set* const* js, const kmsc_set* const* ks, int32_t m,
                          const int64_t* inter_hint, kmsc_set** inter, kmsc_set** j_minus, kmsc_set** k_minus);
/* KmerSet::Add(other) (lib/core/kmer_set.h:164-174) over m sets: the union that
 * KmerSetSet::Get / KmerSetSetReader::Get build (kmer_set_set.h:433-454, 672-755). */
int kmsc_set_union(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t m, kmsc_set** out);
/* KmerSet::Diff (lib/core/kmer_set.h:191-214): |a \ b| + |b \ a|. */
int kmsc_set_diff(kmsc_ctx* ctx, const kmsc_set* a, const kmsc_set* b, int64_t* diff);

/* ---- P1: k-mer counting with cutoff ------------------------------------------------ */
/* Replaces KmerCounter::FromFASTA/FromReads + ToKmerSet
 * (lib/core/kmer_counter.h:64-133, 161-243). fasta = the file bytes (strict
 * 2-line records, '\n' separated). Validation as :163-203: returns KMSC_E_FORMAT
 * with the reference's message. Counts saturate at 255; k-mers with
 * count < cutoff are dropped and counted in *cutoff_count. counts_out (may be
 * NULL) receives a malloc'd uint8 array aligned with the set's key order
 * (free with kmsc_free_host). */
int kmsc_count_fasta(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* fasta, int64_t n_bytes,
                     int canonical, int cutoff, kmsc_set** out, int64_t* cutoff_count,
                     int64_t* n_distinct);
/* Same from reads (one per line, no headers; FromReads semantics). */
int kmsc_count_reads(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* reads, int64_t n_bytes,
                     int canonical, int cutoff, kmsc_set** out, int64_t* cutoff_count,
                     int64_t* n_distinct);
/* KmerCounter::Get (lib/core/kmer_counter.h:246-254) after a counting call with
 * keep_counts: count of one k-mer value in the last counted set of this context. */
int kmsc_count_get(kmsc_ctx* ctx, uint64_t kmer, int* count);
/* uint8 counts of the last counting call, aligned with the key order of the set of
 * ALL distinct k-mers (the set a cutoff <= 1 call returns); out holds n entries. */
int kmsc_count_last_counts(kmsc_ctx* ctx, uint8_t* out, int64_t n);

/* Streaming form for inputs that do not fit one call (config 4: a 20 GB read file): the host
 * feeds chunks made of WHOLE records (FASTA: an even number of lines; reads: whole lines),
 * each chunk is counted on the device and merged into the counter with saturating uint8
 * adds (lib/core/kmer_counter.h:28-38, 105-126); finish applies ToKmerSet(cutoff) (:213-243)
 * and leaves the counter in the context for kmsc_count_get / kmsc_count_last_counts. */
typedef struct kmsc_counter kmsc_counter;
int kmsc_counter_create(kmsc_ctx* ctx, int K, int N, int key_bytes, int canonical, kmsc_counter** out);
int kmsc_counter_add_fasta(kmsc_ctx* ctx, kmsc_counter* c, const char* fasta, int64_t n_bytes);
/* Optional: announce the chunk that will be added NEXT. Its host-to-device copy starts at once on the context's
 * copy stream and overlaps the counting of the chunk added in between; the following kmsc_counter_add_* call
 * with the same pointer and size uses the copy (any other call ignores it). The buffer must stay unchanged
 * until that add call returns; page-locked memory (kmsc_host_alloc_pinned) makes the copy asynchronous. */
int kmsc_counter_prefetch(kmsc_ctx* ctx, kmsc_counter* c, const char* text, int64_t n_bytes);
int kmsc_counter_add_reads(kmsc_ctx* ctx, kmsc_counter* c, const char* reads, int64_t n_bytes);
int kmsc_counter_finish(kmsc_ctx* ctx, kmsc_counter* c, int cutoff, kmsc_set** out, int64_t* cutoff_count,
                        int64_t* n_distinct);
void kmsc_counter_free(kmsc_ctx* ctx, kmsc_counter* c);

/* ---- P5: dense-bitmap Gram for K <= 15 ---------------------------------------------- */
/* out[i*n + j] = |S_i & S_j| over 2^(2K)-bit bitmaps (exact all-bucket matrix);
 * same result as kmsc_pair_counts(bucket_ids = NULL) for duplicate-free sets. */
int kmsc_bitmap_gram(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n, int64_t* out);

/* ---- P6: bucket payload codec (delta + streamvbyte-style 0124) ------------------------ */
/* Internal container (the reference has no on-disk binary format to match:
 * SURVEY.md section 0). encode: device set -> malloc'd host bytes; decode: back. */
int kmsc_codec_encode(kmsc_ctx* ctx, const kmsc_set* set, uint8_t** bytes, int64_t* n_bytes);
int kmsc_codec_decode(kmsc_ctx* ctx, const uint8_t* bytes, int64_t n_bytes, kmsc_set** out);
void kmsc_free_host(void* p);

/* Page-locked host buffers for the streaming inputs: a file reader that fills a
 * buffer obtained here hands kmsc_counter_add_fasta / kmsc_counter_add_reads memory the copy engine reads
 * directly (the same calls from pageable memory are staged by the driver, several times slower). */
int kmsc_host_alloc_pinned(size_t bytes, void** out);
void kmsc_host_free_pinned(void* p);

#ifdef __cplusplus
}
#endif
#endif /* KMSC_H_ */
