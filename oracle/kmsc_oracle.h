/* kmsc_oracle.h -- CPU restatement of kkty/kmer-sets-compression's algorithms for
 * the hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker or the timed CPU baseline. Nothing under
 * kmer-sets-compression_b200/ links, imports or calls it.
 *
 * Parity status: PINNED. tests/test_oracle_golden.py checks every known-answer
 * vector the reference's own tests hold for this path (test/kmer.cc:8-34,
 * test/kmer_counter.cc:12-91, test/kmer_set.cc:72-124, test/spss.cc:13) and
 * tests/golden/ *.json holds outputs of the reference's unmodified headers
 * (oracle/_ref, built by oracle/Makefile from /root/reference/lib) on seeded
 * inputs; tests/test_oracle_vs_ref.py re-runs that comparison live wherever
 * oracle/_ref/libkmsc_ref.so exists. The one un-pinnable item is streamvbyte's
 * byte layout (third-party, not in the tree, never serialised): see
 * kmsc_o_svb0124_*.
 *
 * All k-mers are the reference's 2K-bit values (lib/core/kmer.h:22-46):
 * A=0 C=1 G=2 T=3, first base most significant. "Sets" are ascending arrays of
 * distinct uint64 k-mer values; bucket/key views follow lib/core/kmer_set.h:22-43.
 * Citations are relative to /root/reference/.
 */
#ifndef KMSC_ORACLE_H_
#define KMSC_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- a1-a4: Kmer<K> value type and bucket/key split ---------------------- */
int      kmsc_o_kmer_from_string(const char* s, int K, uint64_t* bits); /* kmer.h:22-46; -1 on non-ACGT */
void     kmsc_o_kmer_to_string(uint64_t bits, int K, char* out);        /* kmer.h:53-81; writes K chars + NUL */
uint64_t kmsc_o_complement(uint64_t bits, int K);                       /* kmer.h:103-129 */
uint64_t kmsc_o_canonical(uint64_t bits, int K);                        /* kmer.h:133 */
uint64_t kmsc_o_next(uint64_t bits, int K, char c);                     /* kmer.h:136-163 */
uint64_t kmsc_o_prev(uint64_t bits, int K, char c);                     /* kmer.h:166-186 */
void     kmsc_o_bucket_key(uint64_t bits, int K, int N, int32_t* bucket, uint64_t* key); /* kmer_set.h:22-31 */
uint64_t kmsc_o_from_bucket_key(int32_t bucket, uint64_t key, int K, int N);             /* kmer_set.h:34-43 */

/* ---- a5-a8: KmerCounter --------------------------------------------------- */
uint8_t  kmsc_o_add_with_max_u8(uint8_t x, uint8_t y);                  /* kmer_counter.h:28-38 */
/* FASTA validation, kmer_counter.h:161-209. lines = n_lines NUL-terminated strings.
 * returns 0 ok, 1 "FASTA files should have an even number of lines", 2 "invalid FASTA file". */
int      kmsc_o_fasta_validate(const char* const* lines, int64_t n_lines);
/* FromReads, kmer_counter.h:64-133: split on 'N', slide K, canonical if asked,
 * uint8 saturating counts. Output: ascending distinct k-mers + counts, malloc'd
 * (free with kmsc_o_free). Returns number of distinct k-mers, <0 on error. */
int64_t  kmsc_o_count_reads(const char* const* reads, int64_t n_reads, int K, int canonical,
                            uint64_t** kmers_out, uint8_t** counts_out);
/* ToKmerSet(cutoff), kmer_counter.h:213-243: keep count >= cutoff (uint8 compare);
 * kept[] must hold n entries; returns #kept, *cutoff_count = #dropped. */
int64_t  kmsc_o_counter_to_set(const uint64_t* kmers, const uint8_t* counts, int64_t n,
                               uint8_t cutoff, uint64_t* kept, int64_t* cutoff_count);

/* ---- a11-a14: KmerSetCompact ---------------------------------------------- */
/* ctor packing, kmer_set_compact.h:206-266: base j of the concatenation goes to
 * bit pair (2j, 2j+1) = (G|T, C|T) of a bit vector. Here the bit vector is
 * little-endian 64-bit words, bit i at word i/64 bit i%64, i.e. word holds 32
 * bases, base j at bits 2(j%32)+{0,1} = {first, second}. words must hold
 * ceil(total_len/32) entries. lengths_minus_k[i] = len_i - K. Returns total bases. */
int64_t  kmsc_o_compact_pack(const char* const* strings, int64_t n, int K,
                             uint64_t* words, uint32_t* lengths_minus_k);
/* ToStrings, kmer_set_compact.h:290-336: unpack string i (len = lengths_minus_k[i]+K)
 * starting at base position pos into out (len chars + NUL). */
void     kmsc_o_compact_unpack(const uint64_t* words, int64_t pos, int64_t len, char* out);
/* Size / Weight, kmer_set_compact.h:90-115 */
int64_t  kmsc_o_compact_size(const uint32_t* lengths_minus_k, int64_t n);
int64_t  kmsc_o_compact_weight(const uint32_t* lengths_minus_k, int64_t n, int K);
/* every k-mer position of every string, in string order (duplicates kept):
 * the common decode loop of kmer_set_compact.h:145-163 and spss.h:1903-1916.
 * out must hold sum(max(0,len-K+1)); returns count. */
int64_t  kmsc_o_spss_kmers(const char* const* strings, int64_t n, int K, int canonical, uint64_t* out);
/* GetSampledKmerSet, kmer_set_compact.h:120-203: out_offs[n_ids+1], keys of
 * bucket bucket_ids[i] ascending at out_keys[out_offs[i]..out_offs[i+1]);
 * duplicates kept; a bucket id listed twice is filled only at its LAST position
 * (map[bucket_ids[i]] = i overwrites, :128-131). out_keys must hold the total
 * k-mer count. Returns number of keys written. */
int64_t  kmsc_o_sampled_set(const char* const* strings, int64_t n, int K, int N, int canonical,
                            const int32_t* bucket_ids, int32_t n_ids,
                            int64_t* out_offs, uint64_t* out_keys);
/* GetKmerSetFromSPSS, spss.h:1861-1941 (hash-set insert = de-duplicate):
 * ascending distinct k-mers into out (capacity = total positions). Returns count. */
int64_t  kmsc_o_set_from_spss(const char* const* strings, int64_t n, int K, int canonical, uint64_t* out);

/* ---- a9-a10: KmerSet algebra on ascending distinct arrays ----------------- */
int64_t  kmsc_o_set_add(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out); /* kmer_set.h:164-174 */
int64_t  kmsc_o_set_sub(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out); /* kmer_set.h:177-187 */
int64_t  kmsc_o_set_intersection(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out); /* kmer_set.h:301-305 */
int64_t  kmsc_o_set_diff(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb);               /* kmer_set.h:191-214 */
uint64_t kmsc_o_set_hash(const uint64_t* a, int64_t n);                                               /* kmer_set.h:224-244 */
/* CSR view: offs[2^N+1] over an ascending k-mer array (bucket = top N bits). */
void     kmsc_o_bucket_offsets(const uint64_t* kmers, int64_t n, int K, int N, int64_t* offs);

/* ---- a16/a18: GetEdgeWeight and the all-pairs matrix ---------------------- */
/* two-pointer merge count over one bucket pair, kmer_set_set.h:165-180
 * (equal => count and advance both, so duplicates count min multiplicity). */
int64_t  kmsc_o_merge_count(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb);
int64_t  kmsc_o_merge_count_u32(const uint32_t* a, int64_t na, const uint32_t* b, int64_t nb);
int64_t  kmsc_o_merge_count_u16(const uint16_t* a, int64_t na, const uint16_t* b, int64_t nb);
/* GetEdgeWeight(i,j), kmer_set_set.h:158-184 over CSR inputs restricted to a
 * bucket list (bucket_ids NULL = all 2^N buckets). key_bytes in {2,4,8}. */
int64_t  kmsc_o_edge_weight(const int64_t* offs_i, const void* keys_i,
                            const int64_t* offs_j, const void* keys_j, int key_bytes,
                            const int32_t* bucket_ids, int32_t n_ids, int32_t n_buckets,
                            int64_t* key_visits);
/* initial all-pairs loop, kmer_set_set.h:187-219: out[n*n] row-major, only i<j
 * written (others 0). n_threads >= 1 (pairs split across pthreads).
 * *key_visits (may be NULL) = sum over pairs and buckets of len_i + len_j. */
void     kmsc_o_pair_counts(const int64_t* const* offs, const void* const* keys, int32_t n_sets,
                            int key_bytes, const int32_t* bucket_ids, int32_t n_ids,
                            int32_t n_buckets, int n_threads, int64_t* out, int64_t* key_visits);

/* ---- a17: greedy driver arithmetic (kmer_set_set.h:267-316) ---------------- */
int32_t  kmsc_o_greedy_interval(int32_t n0);                 /* :267 */
float    kmsc_o_greedy_threshold(int32_t n0);                /* :272-273 */
int      kmsc_o_greedy_should_stop(int64_t total, int64_t updated, int32_t n0); /* :287-297 */
/* argmax with the adopted tie-break (max weight, then smallest (j,k)); w is a
 * dense n*n row-major matrix, only i<j read. Returns weight, 0 => stop (:319). */
int64_t  kmsc_o_greedy_argmax(const int64_t* w, int32_t n, int32_t* j, int32_t* k);

/* ---- a22: ParallelDisjointSet, serial semantics (parallel_disjoint_set.h) -- */
typedef struct kmsc_o_dsu kmsc_o_dsu;
kmsc_o_dsu* kmsc_o_dsu_new(int32_t n);
void     kmsc_o_dsu_free(kmsc_o_dsu*);
int32_t  kmsc_o_dsu_find(kmsc_o_dsu*, int32_t x);            /* :24-40 */
int      kmsc_o_dsu_same(kmsc_o_dsu*, int32_t x, int32_t y); /* :43-50 */
void     kmsc_o_dsu_unite(kmsc_o_dsu*, int32_t x, int32_t y);/* :53-78 */

/* bucket sizes of GetSampledKmerSet over all buckets (kmer_set_compact.h:145-163) for a sequence
 * given as codes 0..3, without building the sets: hist[2^N] += k-mer positions per bucket. */
void     kmsc_o_bucket_histogram(const uint8_t* codes, int64_t n, int K, int N, int canonical, int64_t* hist);

/* ---- `mst` driver (north_star; no counterpart in the snapshot: SURVEY.md App. C) ---------- */
/* w = exact n x n intersection matrix with the set sizes on the diagonal. Kruskal over
 * d(i,j) = |S_i| + |S_j| - 2 w[i][j], candidates ordered by (d, i, j), union-find as
 * parallel_disjoint_set.h:53-78; tree oriented breadth-first from node 0, neighbours in
 * ascending index order. edges: (parent, child) x (n-1); dist: n-1. Returns the edge count. */
int32_t  kmsc_o_mst(const int64_t* w, int32_t n, int32_t* edges, int64_t* dist);

/* ---- streamvbyte "0124" (third party, v0.4.1; published format) ------------ */
size_t   kmsc_o_svb0124_max_bytes(uint32_t n);
size_t   kmsc_o_svb0124_encode(const uint32_t* in, uint32_t n, uint8_t* out);
size_t   kmsc_o_svb0124_decode(const uint8_t* in, uint32_t* out, uint32_t n);

/* ---- P6: "KMSC" container (delta + 1/2/3/4-byte codes); no reference counterpart ---- */
uint8_t* kmsc_o_codec_encode(int K, int N, int key_bytes, const int64_t* offs, const void* keys, int64_t* n_bytes);
int      kmsc_o_codec_decode(const uint8_t* bytes, int64_t n_bytes, int* K, int* N, int* key_bytes, int64_t* n_keys,
                             int64_t* offs, void* keys);

void     kmsc_o_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
