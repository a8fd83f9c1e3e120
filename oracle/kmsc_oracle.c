/* kmsc_oracle.c -- CPU restatement of the reference's hot-path algorithms.
 * TEST INFRASTRUCTURE ONLY: see kmsc_oracle.h for the rules and parity status.
 * Each function cites the reference file:line it follows (relative to
 * /root/reference/). Plain C99 + pthreads. */
#include "kmsc_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>


/* ------------------------------------------------------------------------- */
/* P6: "KMSC" container = delta + streamvbyte-style 1/2/3/4-byte codes        */
/* ------------------------------------------------------------------------- */
/* No counterpart in the reference (it never serialises a binary k-mer set); the byte-code
 * scheme is the one KmerSetCompact applies to its string lengths
 * (lib/core/kmer_set_compact.h:256-266: control bytes first, 2-bit codes LSB first,
 * little-endian data bytes), here with the classic 1/2/3/4-byte code lengths. Layout in
 * kmer-sets-compression_b200/csrc/codec.cu. Checks the GPU codec byte for byte. */

static uint32_t codec_code(uint32_t v) { return v < (1u << 8) ? 0u : v < (1u << 16) ? 1u : v < (1u << 24) ? 2u : 3u; }

/* values -> ctrl (ceil(m/4) bytes) + data; returns data bytes */
static size_t codec_stream_encode(const uint32_t* v, uint64_t m, uint8_t* ctrl, uint8_t* data) {
  size_t w = 0;
  memset(ctrl, 0, (size_t)((m + 3) / 4));
  for (uint64_t i = 0; i < m; i++) {
    const uint32_t c = codec_code(v[i]);
    ctrl[i >> 2] |= (uint8_t)(c << (2 * (i & 3)));
    for (uint32_t t = 0; t <= c; t++) data[w++] = (uint8_t)(v[i] >> (8 * t));
  }
  return w;
}

static int codec_stream_decode(const uint8_t* ctrl, const uint8_t* data, size_t n_data, uint64_t m, uint32_t* v) {
  size_t r = 0;
  for (uint64_t i = 0; i < m; i++) {
    const uint32_t len = ((ctrl[i >> 2] >> (2 * (i & 3))) & 3u) + 1;
    if (r + len > n_data) return -1;
    uint32_t x = 0;
    for (uint32_t t = 0; t < len; t++) x |= (uint32_t)data[r + t] << (8 * t);
    v[i] = x;
    r += len;
  }
  return r == n_data ? 0 : -1;
}

/* CSR (offs int64[2^N + 1], keys of key_bytes each) -> malloc'd container */
uint8_t* kmsc_o_codec_encode(int K, int N, int key_bytes, const int64_t* offs, const void* keys, int64_t* n_bytes) {
  const uint64_t nb = (uint64_t)1 << N;
  const uint64_t n = (uint64_t)offs[nb];
  const int wpk = key_bytes == 8 ? 2 : 1;
  const uint64_t m = n * (uint64_t)wpk;
  uint32_t* sv = (uint32_t*)malloc((size_t)(nb + 1) * 4);
  uint32_t* kv = (uint32_t*)malloc((size_t)(m + 1) * 4);
  for (uint64_t b = 0; b < nb; b++) sv[b] = (uint32_t)(offs[b + 1] - offs[b]);
  for (uint64_t b = 0; b < nb; b++) {
    uint64_t prev = 0;
    for (int64_t i = offs[b]; i < offs[b + 1]; i++) {
      uint64_t k = key_bytes == 2 ? ((const uint16_t*)keys)[i] : key_bytes == 4 ? ((const uint32_t*)keys)[i] : ((const uint64_t*)keys)[i];
      const uint64_t d = k - prev;
      prev = k;
      kv[(uint64_t)i * wpk] = (uint32_t)d;
      if (wpk == 2) kv[(uint64_t)i * wpk + 1] = (uint32_t)(d >> 32);
    }
  }
  const size_t cap = 48 + (size_t)(nb + 3) / 4 + (size_t)nb * 4 + (size_t)(m + 3) / 4 + (size_t)m * 4 + 16;
  uint8_t* out = (uint8_t*)malloc(cap);
  size_t w = 48;
  const size_t n_sctrl = (size_t)(nb + 3) / 4;
  const size_t n_sdata = codec_stream_encode(sv, nb, out + w, out + w + n_sctrl);
  w += n_sctrl + n_sdata;
  const size_t n_kctrl = (size_t)(m + 3) / 4;
  const size_t n_kdata = codec_stream_encode(kv, m, out + w, out + w + n_kctrl);
  w += n_kctrl + n_kdata;
  const uint32_t h32[6] = {0x43534D4Bu, 1u, (uint32_t)K, (uint32_t)N, (uint32_t)key_bytes, (uint32_t)wpk};
  const uint64_t h64[3] = {n, (uint64_t)n_sdata, (uint64_t)n_kdata};
  memcpy(out, h32, 24);
  memcpy(out + 24, h64, 24);
  free(sv); free(kv);
  *n_bytes = (int64_t)w;
  return out;
}

/* container -> header fields; offs (int64[2^N + 1]) and keys (caller-sized by a first call
 * with offs == NULL, which only fills the header outputs). Returns 0, -1 on corruption. */
int kmsc_o_codec_decode(const uint8_t* bytes, int64_t n_bytes, int* K, int* N, int* key_bytes, int64_t* n_keys,
                        int64_t* offs, void* keys) {
  if (n_bytes < 48) return -1;
  uint32_t h32[6]; uint64_t h64[3];
  memcpy(h32, bytes, 24); memcpy(h64, bytes + 24, 24);
  if (h32[0] != 0x43534D4Bu || h32[1] != 1u) return -1;
  *K = (int)h32[2]; *N = (int)h32[3]; *key_bytes = (int)h32[4]; *n_keys = (int64_t)h64[0];
  if (!offs) return 0;
  const uint64_t nb = (uint64_t)1 << *N;
  const int wpk = (int)h32[5];
  const uint64_t n = h64[0], m = n * (uint64_t)wpk;
  const size_t n_sctrl = (size_t)(nb + 3) / 4, n_kctrl = (size_t)(m + 3) / 4;
  if ((uint64_t)n_bytes != 48 + n_sctrl + h64[1] + n_kctrl + h64[2]) return -1;
  uint32_t* sv = (uint32_t*)malloc((size_t)(nb + 1) * 4);
  uint32_t* kv = (uint32_t*)malloc((size_t)(m + 1) * 4);
  const uint8_t* p = bytes + 48;
  int rc = codec_stream_decode(p, p + n_sctrl, (size_t)h64[1], nb, sv);
  p += n_sctrl + h64[1];
  if (rc == 0) rc = codec_stream_decode(p, p + n_kctrl, (size_t)h64[2], m, kv);
  if (rc == 0) {
    offs[0] = 0;
    for (uint64_t b = 0; b < nb; b++) offs[b + 1] = offs[b] + sv[b];
    if ((uint64_t)offs[nb] != n) rc = -1;
  }
  if (rc == 0) {
    for (uint64_t b = 0; b < nb; b++) {
      uint64_t prev = 0;
      for (int64_t i = offs[b]; i < offs[b + 1]; i++) {
        uint64_t d = kv[(uint64_t)i * wpk];
        if (wpk == 2) d |= (uint64_t)kv[(uint64_t)i * wpk + 1] << 32;
        prev += d;
        if (*key_bytes == 2) ((uint16_t*)keys)[i] = (uint16_t)prev;
        else if (*key_bytes == 4) ((uint32_t*)keys)[i] = (uint32_t)prev;
        else ((uint64_t*)keys)[i] = prev;
      }
    }
  }
  free(sv); free(kv);
  return rc;
}

void kmsc_o_free(void* p) { free(p); }

/* ------------------------------------------------------------------------- */
/* a1-a4: Kmer<K>                                                             */
/* ------------------------------------------------------------------------- */

static int base_code(char c) {
  switch (c) { /* lib/core/kmer.h:28-41 */
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

int kmsc_o_kmer_from_string(const char* s, int K, uint64_t* bits) {
  uint64_t b = 0; /* lib/core/kmer.h:22-46: bits <<= 2; bits += code */
  for (int i = 0; i < K; i++) {
    int c = base_code(s[i]);
    if (c < 0) return -1;
    b = (b << 2) + (uint64_t)c;
  }
  *bits = b;
  return 0;
}

void kmsc_o_kmer_to_string(uint64_t bits, int K, char* out) {
  static const char L[4] = {'A', 'C', 'G', 'T'}; /* lib/core/kmer.h:53-81 */
  for (int i = 0; i < K; i++) {
    out[K - 1 - i] = L[bits & 3];
    bits >>= 2;
  }
  out[K] = '\0';
}

uint64_t kmsc_o_complement(uint64_t bits, int K) {
  uint64_t c = 0; /* lib/core/kmer.h:103-129: reverse order, code -> 3 - code */
  for (int i = 0; i < K; i++) {
    c = (c << 2) + (3 - (bits & 3));
    bits >>= 2;
  }
  return c;
}

uint64_t kmsc_o_canonical(uint64_t bits, int K) {
  uint64_t c = kmsc_o_complement(bits, K); /* lib/core/kmer.h:133, :232-235 */
  return c < bits ? c : bits;
}

uint64_t kmsc_o_next(uint64_t bits, int K, char c) {
  uint64_t mask = ~(uint64_t)0 >> (64 - K * 2); /* lib/core/kmer.h:136-163 */
  return ((bits << 2) & mask) + (uint64_t)base_code(c);
}

uint64_t kmsc_o_prev(uint64_t bits, int K, char c) {
  return (bits >> 2) + ((uint64_t)base_code(c) << ((K - 1) * 2)); /* lib/core/kmer.h:166-186 */
}

void kmsc_o_bucket_key(uint64_t bits, int K, int N, int32_t* bucket, uint64_t* key) {
  const int n_key_bits = K * 2 - N; /* lib/core/kmer_set.h:22-31 */
  *bucket = (int32_t)(bits >> n_key_bits);
  *key = bits % ((uint64_t)1 << n_key_bits);
}

uint64_t kmsc_o_from_bucket_key(int32_t bucket, uint64_t key, int K, int N) {
  const int n_key_bits = K * 2 - N; /* lib/core/kmer_set.h:34-43 */
  return ((uint64_t)bucket << n_key_bits) + key;
}

/* ------------------------------------------------------------------------- */
/* helpers: LSD radix sort of uint64                                          */
/* ------------------------------------------------------------------------- */

static void sort_u64(uint64_t* a, int64_t n) {
  if (n < 2) return;
  uint64_t* tmp = (uint64_t*)malloc((size_t)n * sizeof(uint64_t));
  uint64_t* src = a;
  uint64_t* dst = tmp;
  uint64_t orall = 0;
  for (int64_t i = 0; i < n; i++) orall |= a[i];
  for (int shift = 0; shift < 64; shift += 8) {
    if ((orall >> shift) == 0) break; /* no higher digits left */
    int64_t hist[257];
    memset(hist, 0, sizeof(hist));
    for (int64_t i = 0; i < n; i++) hist[((src[i] >> shift) & 0xff) + 1]++;
    for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
    for (int64_t i = 0; i < n; i++) dst[hist[(src[i] >> shift) & 0xff]++] = src[i];
    uint64_t* t = src; src = dst; dst = t;
  }
  if (src != a) memcpy(a, src, (size_t)n * sizeof(uint64_t));
  free(tmp);
}

static int64_t unique_u64(uint64_t* a, int64_t n) {
  if (n == 0) return 0;
  int64_t m = 1;
  for (int64_t i = 1; i < n; i++)
    if (a[i] != a[m - 1]) a[m++] = a[i];
  return m;
}

/* ------------------------------------------------------------------------- */
/* a5-a8: KmerCounter                                                         */
/* ------------------------------------------------------------------------- */

uint8_t kmsc_o_add_with_max_u8(uint8_t x, uint8_t y) {
  int64_t s = (int64_t)x + (int64_t)y; /* lib/core/kmer_counter.h:28-38 */
  return (uint8_t)(s < 255 ? s : 255);
}

int kmsc_o_fasta_validate(const char* const* lines, int64_t n_lines) {
  if (n_lines % 2 != 0) return 1; /* lib/core/kmer_counter.h:163-166 */
  for (int64_t i = 0; i < n_lines; i++) {
    const char* l = lines[i];
    if (i % 2 == 0) {
      if (l[0] == '\0' || l[0] != '>') return 2; /* :179-183 */
    } else {
      for (const char* p = l; *p; p++) /* :186-191 */
        if (*p != 'A' && *p != 'C' && *p != 'G' && *p != 'T' && *p != 'N') return 2;
    }
  }
  return 0;
}

/* walks every K-window that holds no 'N' (lib/core/kmer_counter.h:78-92: the
 * read is split on 'N' and each fragment slid separately) and calls emit. */
static int64_t for_each_read_kmer(const char* read, int K, int canonical, uint64_t* out) {
  int64_t n = 0;
  const char* frag = read;
  for (;;) {
    const char* e = frag;
    while (*e && *e != 'N') e++;
    int64_t len = e - frag;
    for (int64_t j = 0; j + K <= len; j++) {
      uint64_t b;
      if (kmsc_o_kmer_from_string(frag + j, K, &b) != 0) return -1;
      if (out) out[n] = canonical ? kmsc_o_canonical(b, K) : b;
      n++;
    }
    if (!*e) break;
    frag = e + 1;
  }
  return n;
}

int64_t kmsc_o_count_reads(const char* const* reads, int64_t n_reads, int K, int canonical,
                           uint64_t** kmers_out, uint8_t** counts_out) {
  int64_t total = 0;
  for (int64_t i = 0; i < n_reads; i++) {
    int64_t c = for_each_read_kmer(reads[i], K, canonical, NULL);
    if (c < 0) return -1;
    total += c;
  }
  uint64_t* all = (uint64_t*)malloc((size_t)(total ? total : 1) * sizeof(uint64_t));
  int64_t pos = 0;
  for (int64_t i = 0; i < n_reads; i++) pos += for_each_read_kmer(reads[i], K, canonical, all + pos);
  sort_u64(all, total);
  uint8_t* counts = (uint8_t*)malloc((size_t)(total ? total : 1));
  int64_t m = 0;
  for (int64_t i = 0; i < total;) { /* counts saturate: kmer_counter.h:94, :117 */
    int64_t j = i;
    uint8_t c = 0;
    while (j < total && all[j] == all[i]) { c = kmsc_o_add_with_max_u8(c, 1); j++; }
    all[m] = all[i];
    counts[m] = c;
    m++;
    i = j;
  }
  *kmers_out = all;
  *counts_out = counts;
  return m;
}

int64_t kmsc_o_counter_to_set(const uint64_t* kmers, const uint8_t* counts, int64_t n,
                              uint8_t cutoff, uint64_t* kept, int64_t* cutoff_count) {
  int64_t m = 0, cut = 0; /* lib/core/kmer_counter.h:222-236 */
  for (int64_t i = 0; i < n; i++) {
    if (counts[i] < cutoff) { cut++; continue; }
    kept[m++] = kmers[i];
  }
  *cutoff_count = cut;
  return m;
}

/* ------------------------------------------------------------------------- */
/* a11-a14: KmerSetCompact                                                    */
/* ------------------------------------------------------------------------- */

int64_t kmsc_o_compact_pack(const char* const* strings, int64_t n, int K,
                            uint64_t* words, uint32_t* lengths_minus_k) {
  int64_t pos = 0; /* lib/core/kmer_set_compact.h:206-255 */
  for (int64_t i = 0; i < n; i++) {
    int64_t len = (int64_t)strlen(strings[i]);
    lengths_minus_k[i] = (uint32_t)(len - K);
    pos += len;
  }
  memset(words, 0, (size_t)((pos + 31) / 32) * sizeof(uint64_t));
  pos = 0;
  for (int64_t i = 0; i < n; i++) {
    const char* s = strings[i];
    for (int64_t j = 0; s[j]; j++, pos++) {
      int first = (s[j] == 'G' || s[j] == 'T');  /* data_[position + j*2]     */
      int second = (s[j] == 'C' || s[j] == 'T'); /* data_[position + j*2 + 1] */
      words[pos / 32] |= (uint64_t)(first | (second << 1)) << (2 * (pos % 32));
    }
  }
  return pos;
}

void kmsc_o_compact_unpack(const uint64_t* words, int64_t pos, int64_t len, char* out) {
  for (int64_t j = 0; j < len; j++, pos++) { /* lib/core/kmer_set_compact.h:309-327 */
    unsigned v = (unsigned)(words[pos / 32] >> (2 * (pos % 32))) & 3u;
    int first = v & 1, second = (v >> 1) & 1;
    out[j] = first ? (second ? 'T' : 'G') : (second ? 'C' : 'A');
  }
  out[len] = '\0';
}

int64_t kmsc_o_compact_size(const uint32_t* lengths_minus_k, int64_t n) {
  int64_t s = 0; /* lib/core/kmer_set_compact.h:90-112: sum(len - K + 1) */
  for (int64_t i = 0; i < n; i++) s += (int64_t)lengths_minus_k[i] + 1;
  return s;
}

int64_t kmsc_o_compact_weight(const uint32_t* lengths_minus_k, int64_t n, int K) {
  int64_t s = 0; /* lib/core/kmer_set_compact.h:115: data_.size() / 2 = sum(len) */
  for (int64_t i = 0; i < n; i++) s += (int64_t)lengths_minus_k[i] + K;
  return s;
}

int64_t kmsc_o_spss_kmers(const char* const* strings, int64_t n, int K, int canonical, uint64_t* out) {
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++) {
    const char* s = strings[i];
    int64_t len = (int64_t)strlen(s);
    for (int64_t j = 0; j < len - K + 1; j++) { /* kmer_set_compact.h:149-151 */
      uint64_t b;
      if (kmsc_o_kmer_from_string(s + j, K, &b) != 0) return -1;
      out[m++] = canonical ? kmsc_o_canonical(b, K) : b;
    }
  }
  return m;
}

int64_t kmsc_o_sampled_set(const char* const* strings, int64_t n, int K, int N, int canonical,
                           const int32_t* bucket_ids, int32_t n_ids,
                           int64_t* out_offs, uint64_t* out_keys) {
  const int32_t nb = (int32_t)1 << N;
  int32_t* map = (int32_t*)malloc((size_t)nb * sizeof(int32_t)); /* :127-131 */
  for (int32_t b = 0; b < nb; b++) map[b] = -1;
  for (int32_t i = 0; i < n_ids; i++) map[bucket_ids[i]] = i;
  int64_t total = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t len = (int64_t)strlen(strings[i]);
    if (len - K + 1 > 0) total += len - K + 1;
  }
  uint64_t* all = (uint64_t*)malloc((size_t)(total ? total : 1) * sizeof(uint64_t));
  int64_t m = kmsc_o_spss_kmers(strings, n, K, canonical, all);
  if (m < 0) { free(all); free(map); return -1; }
  /* count per selected position, then fill, then sort each (:190-200) */
  for (int32_t i = 0; i <= n_ids; i++) out_offs[i] = 0;
  for (int64_t t = 0; t < m; t++) {
    int32_t b; uint64_t key;
    kmsc_o_bucket_key(all[t], K, N, &b, &key);
    if (map[b] >= 0) out_offs[map[b] + 1]++;
  }
  for (int32_t i = 0; i < n_ids; i++) out_offs[i + 1] += out_offs[i];
  int64_t* cur = (int64_t*)malloc((size_t)(n_ids + 1) * sizeof(int64_t));
  memcpy(cur, out_offs, (size_t)(n_ids + 1) * sizeof(int64_t));
  for (int64_t t = 0; t < m; t++) {
    int32_t b; uint64_t key;
    kmsc_o_bucket_key(all[t], K, N, &b, &key);
    if (map[b] >= 0) out_keys[cur[map[b]]++] = key;
  }
  for (int32_t i = 0; i < n_ids; i++) sort_u64(out_keys + out_offs[i], out_offs[i + 1] - out_offs[i]);
  int64_t written = out_offs[n_ids];
  free(cur); free(all); free(map);
  return written;
}

int64_t kmsc_o_set_from_spss(const char* const* strings, int64_t n, int K, int canonical, uint64_t* out) {
  int64_t m = kmsc_o_spss_kmers(strings, n, K, canonical, out); /* spss.h:1903-1916 */
  if (m < 0) return -1;
  sort_u64(out, m);
  return unique_u64(out, m); /* hash-set insert, spss.h:1925 -> kmer_set.h:77-83 */
}

/* ------------------------------------------------------------------------- */
/* a9-a10: KmerSet algebra                                                    */
/* ------------------------------------------------------------------------- */

int64_t kmsc_o_set_add(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out) {
  int64_t i = 0, j = 0, m = 0; /* lib/core/kmer_set.h:164-174: insert every key of other */
  while (i < na && j < nb) {
    if (a[i] < b[j]) out[m++] = a[i++];
    else if (b[j] < a[i]) out[m++] = b[j++];
    else { out[m++] = a[i]; i++; j++; }
  }
  while (i < na) out[m++] = a[i++];
  while (j < nb) out[m++] = b[j++];
  return m;
}

int64_t kmsc_o_set_sub(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out) {
  int64_t i = 0, j = 0, m = 0; /* lib/core/kmer_set.h:177-187: erase every key of other */
  while (i < na) {
    while (j < nb && b[j] < a[i]) j++;
    if (j < nb && b[j] == a[i]) { i++; continue; }
    out[m++] = a[i++];
  }
  return m;
}

int64_t kmsc_o_set_intersection(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out) {
  /* lib/core/kmer_set.h:301-305: lhs.Sub(Sub(lhs, rhs)) */
  uint64_t* d = (uint64_t*)malloc((size_t)(na ? na : 1) * sizeof(uint64_t));
  int64_t nd = kmsc_o_set_sub(a, na, b, nb, d);
  int64_t m = kmsc_o_set_sub(a, na, d, nd, out);
  free(d);
  return m;
}

int64_t kmsc_o_set_diff(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb) {
  int64_t i = 0, j = 0, c = 0; /* lib/core/kmer_set.h:191-214: |b\a| + |a\b| */
  while (i < na && j < nb) {
    if (a[i] < b[j]) { c++; i++; }
    else if (b[j] < a[i]) { c++; j++; }
    else { i++; j++; }
  }
  return c + (na - i) + (nb - j);
}

uint64_t kmsc_o_set_hash(const uint64_t* a, int64_t n) {
  uint64_t h = 0; /* lib/core/kmer_set.h:224-244 with kmer.h:211 (Hash() = bits) */
  for (int64_t i = 0; i < n; i++) h ^= a[i];
  return h;
}

void kmsc_o_bucket_offsets(const uint64_t* kmers, int64_t n, int K, int N, int64_t* offs) {
  const int64_t nb = (int64_t)1 << N;
  const int shift = 2 * K - N;
  int64_t p = 0;
  for (int64_t b = 0; b <= nb; b++) {
    while (p < n && (int64_t)(kmers[p] >> shift) < b) p++;
    offs[b] = p;
  }
}

/* ------------------------------------------------------------------------- */
/* a16/a18: GetEdgeWeight                                                     */
/* ------------------------------------------------------------------------- */

#define MERGE_COUNT_BODY                                                      \
  int64_t i = 0, j = 0, c = 0; /* lib/core/kmer_set_set.h:165-180 */          \
  while (i < na && j < nb) {                                                  \
    if (a[i] < b[j]) i++;                                                     \
    else if (a[i] > b[j]) j++;                                                \
    else { c++; i++; j++; }                                                   \
  }                                                                           \
  return c;

int64_t kmsc_o_merge_count(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb) { MERGE_COUNT_BODY }
int64_t kmsc_o_merge_count_u32(const uint32_t* a, int64_t na, const uint32_t* b, int64_t nb) { MERGE_COUNT_BODY }
int64_t kmsc_o_merge_count_u16(const uint16_t* a, int64_t na, const uint16_t* b, int64_t nb) { MERGE_COUNT_BODY }

int64_t kmsc_o_edge_weight(const int64_t* offs_i, const void* keys_i,
                           const int64_t* offs_j, const void* keys_j, int key_bytes,
                           const int32_t* bucket_ids, int32_t n_ids, int32_t n_buckets,
                           int64_t* key_visits) {
  int64_t count = 0, visits = 0; /* lib/core/kmer_set_set.h:158-184 */
  const int32_t nb = bucket_ids ? n_ids : n_buckets;
  /* a bucket id listed twice fills only ONE position of the sampled set
   * (map[bucket_ids[i]] = i, kmer_set_compact.h:127-131; the other position stays
   * empty), so it is merged once */
  uint8_t* seen = bucket_ids ? (uint8_t*)calloc((size_t)n_buckets, 1) : NULL;
  for (int32_t t = 0; t < nb; t++) {
    const int32_t b = bucket_ids ? bucket_ids[t] : t;
    if (seen) { if (seen[b]) continue; seen[b] = 1; }
    const int64_t ai = offs_i[b], na = offs_i[b + 1] - ai;
    const int64_t bj = offs_j[b], nbk = offs_j[b + 1] - bj;
    visits += na + nbk;
    switch (key_bytes) {
      case 2: count += kmsc_o_merge_count_u16((const uint16_t*)keys_i + ai, na, (const uint16_t*)keys_j + bj, nbk); break;
      case 4: count += kmsc_o_merge_count_u32((const uint32_t*)keys_i + ai, na, (const uint32_t*)keys_j + bj, nbk); break;
      default: count += kmsc_o_merge_count((const uint64_t*)keys_i + ai, na, (const uint64_t*)keys_j + bj, nbk); break;
    }
  }
  free(seen);
  if (key_visits) *key_visits = visits;
  return count;
}

typedef struct {
  const int64_t* const* offs;
  const void* const* keys;
  int32_t n_sets;
  int key_bytes;
  const int32_t* bucket_ids;
  int32_t n_ids, n_buckets;
  int64_t* out;
  int64_t pair_begin, pair_end;
  int64_t visits;
} pc_task;

static void pair_from_index(int64_t p, int32_t n, int32_t* i, int32_t* j) {
  int32_t a = 0; /* pairs enumerated as lib/core/kmer_set_set.h:199-203 */
  int64_t row = n - 1;
  while (p >= row) { p -= row; a++; row--; }
  *i = a;
  *j = a + 1 + (int32_t)p;
}

static void* pc_worker(void* arg) {
  pc_task* t = (pc_task*)arg;
  for (int64_t p = t->pair_begin; p < t->pair_end; p++) {
    int32_t i, j;
    pair_from_index(p, t->n_sets, &i, &j);
    int64_t v = 0;
    int64_t w = kmsc_o_edge_weight(t->offs[i], t->keys[i], t->offs[j], t->keys[j], t->key_bytes,
                                   t->bucket_ids, t->n_ids, t->n_buckets, &v);
    t->out[(int64_t)i * t->n_sets + j] = w;
    t->visits += v;
  }
  return NULL;
}

void kmsc_o_pair_counts(const int64_t* const* offs, const void* const* keys, int32_t n_sets,
                        int key_bytes, const int32_t* bucket_ids, int32_t n_ids,
                        int32_t n_buckets, int n_threads, int64_t* out, int64_t* key_visits) {
  const int64_t n_pairs = (int64_t)n_sets * (n_sets - 1) / 2;
  memset(out, 0, (size_t)n_sets * n_sets * sizeof(int64_t));
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_pairs && n_pairs > 0) n_threads = (int)n_pairs;
  pc_task* tasks = (pc_task*)calloc((size_t)n_threads, sizeof(pc_task));
  pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  for (int t = 0; t < n_threads; t++) {
    tasks[t].offs = offs; tasks[t].keys = keys; tasks[t].n_sets = n_sets;
    tasks[t].key_bytes = key_bytes; tasks[t].bucket_ids = bucket_ids;
    tasks[t].n_ids = n_ids; tasks[t].n_buckets = n_buckets; tasks[t].out = out;
    tasks[t].pair_begin = n_pairs * t / n_threads;
    tasks[t].pair_end = n_pairs * (t + 1) / n_threads;
    if (n_threads > 1) pthread_create(&th[t], NULL, pc_worker, &tasks[t]);
    else pc_worker(&tasks[t]);
  }
  int64_t v = 0;
  for (int t = 0; t < n_threads; t++) {
    if (n_threads > 1) pthread_join(th[t], NULL);
    v += tasks[t].visits;
  }
  if (key_visits) *key_visits = v;
  free(tasks); free(th);
}

/* ------------------------------------------------------------------------- */
/* a17: greedy driver arithmetic                                              */
/* ------------------------------------------------------------------------- */

int32_t kmsc_o_greedy_interval(int32_t n0) { return n0 / 8 + 1; } /* kmer_set_set.h:267 */

float kmsc_o_greedy_threshold(int32_t n0) {
  /* kmer_set_set.h:272-273: double arithmetic, narrowed to float */
  const int interval = kmsc_o_greedy_interval(n0);
  const float t = 0.1 * interval / (size_t)n0;
  return t;
}

int kmsc_o_greedy_should_stop(int64_t total, int64_t updated, int32_t n0) {
  /* kmer_set_set.h:287-297: float(total - updated) / total (int64 -> float promotion) */
  const float improvement = (float)(total - updated) / total;
  return improvement <= kmsc_o_greedy_threshold(n0);
}

int64_t kmsc_o_greedy_argmax(const int64_t* w, int32_t n, int32_t* j, int32_t* k) {
  int64_t best = 0; /* kmer_set_set.h:308-316, strict '>' in ascending (j,k) order */
  *j = -1; *k = -1;
  for (int32_t a = 0; a < n; a++)
    for (int32_t b = a + 1; b < n; b++)
      if (w[(int64_t)a * n + b] > best) { best = w[(int64_t)a * n + b]; *j = a; *k = b; }
  return best;
}

/* ------------------------------------------------------------------------- */
/* a22: ParallelDisjointSet (serial semantics)                                */
/* ------------------------------------------------------------------------- */

struct kmsc_o_dsu { int32_t n; uint64_t* a; }; /* word = rank<<32 | parent, :109-110 */

static int32_t dsu_rank(const kmsc_o_dsu* d, int32_t i) { return (int32_t)(d->a[i] >> 32); }
static int32_t dsu_next(const kmsc_o_dsu* d, int32_t i) { return (int32_t)(d->a[i] & 0xffffffffu); }
static int dsu_less(const kmsc_o_dsu* d, int32_t x, int32_t y) { /* :98-106 */
  int32_t rx = dsu_rank(d, x), ry = dsu_rank(d, y);
  if (rx < ry) return 1;
  if (rx > ry) return 0;
  return x < y;
}

kmsc_o_dsu* kmsc_o_dsu_new(int32_t n) {
  kmsc_o_dsu* d = (kmsc_o_dsu*)malloc(sizeof(kmsc_o_dsu));
  d->n = n;
  d->a = (uint64_t*)malloc((size_t)(n ? n : 1) * sizeof(uint64_t));
  for (int32_t i = 0; i < n; i++) d->a[i] = (uint64_t)i; /* :17-21 */
  return d;
}

void kmsc_o_dsu_free(kmsc_o_dsu* d) { if (d) { free(d->a); free(d); } }

int32_t kmsc_o_dsu_find(kmsc_o_dsu* d, int32_t x) {
  int32_t y = x; /* :24-40 */
  while (x != dsu_next(d, x)) x = dsu_next(d, x);
  while (dsu_less(d, y, x)) {
    int32_t nxt = dsu_next(d, y);
    d->a[y] = ((uint64_t)dsu_rank(d, y) << 32) + (uint64_t)x;
    y = nxt;
  }
  return x;
}

int kmsc_o_dsu_same(kmsc_o_dsu* d, int32_t x, int32_t y) {
  return kmsc_o_dsu_find(d, x) == kmsc_o_dsu_find(d, y); /* :43-50 */
}

void kmsc_o_dsu_unite(kmsc_o_dsu* d, int32_t x, int32_t y) {
  x = kmsc_o_dsu_find(d, x); /* :53-78 */
  y = kmsc_o_dsu_find(d, y);
  if (x == y) return;
  int32_t rx = dsu_rank(d, x), ry = dsu_rank(d, y);
  if (rx > ry || (rx == ry && x > y)) {
    int32_t t = x; x = y; y = t;
    t = rx; rx = ry; ry = t;
  }
  d->a[x] = ((uint64_t)rx << 32) + (uint64_t)y;          /* UpdateRoot(x, rx, y, rx) */
  if (rx == ry) d->a[y] = ((uint64_t)(ry + 1) << 32) + (uint64_t)y; /* rank bump, :73-75 */
}

/* k-mer positions per bucket (duplicates counted) of 2-bit codes 0..3: the bucket sizes
 * GetSampledKmerSet would produce (kmer_set_compact.h:145-163), without building the sets.
 * Used by bench.py to state the key-visits of a reference run. hist: 2^N entries, added to. */
void kmsc_o_bucket_histogram(const uint8_t* codes, int64_t n, int K, int N, int canonical, int64_t* hist) {
  if (n < K) return;
  const uint64_t mask = K == 32 ? ~(uint64_t)0 : (((uint64_t)1 << (2 * K)) - 1);
  uint64_t fwd = 0, rc = 0;
  for (int64_t i = 0; i < n; i++) {
    const uint64_t c = codes[i] & 3u;
    fwd = ((fwd << 2) | c) & mask;
    rc = (rc >> 2) | ((3 - c) << (2 * (K - 1)));
    if (i >= K - 1) {
      const uint64_t v = (canonical && rc < fwd) ? rc : fwd;
      hist[v >> (2 * K - N)]++;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* `mst` driver (north_star variant; the snapshot has no counterpart) */
/* ------------------------------------------------------------------------- */

typedef struct { int64_t d; int32_t i, j; } mst_cand;

static int mst_cand_cmp(const void* a, const void* b) {
  const mst_cand* x = (const mst_cand*)a;
  const mst_cand* y = (const mst_cand*)b;
  if (x->d != y->d) return x->d < y->d ? -1 : 1;
  if (x->i != y->i) return x->i < y->i ? -1 : 1;
  if (x->j != y->j) return x->j < y->j ? -1 : 1;
  return 0;
}

/* w: exact n x n intersection matrix (diagonal = set sizes, upper triangle read).
 * d(i,j) = |S_i| + |S_j| - 2 w[i][j]; candidate edges in (d, i, j) ascending order; Kruskal
 * with the union-find of lib/core/parallel_disjoint_set.h:53-78, 98-106 (above); the tree is
 * oriented by a breadth-first walk from node 0 that visits a node's neighbours in ascending
 * index order. Output: n - 1 rows (parent, child) in visiting order and their distances.
 * Returns the number of edges (n - 1, or less if some d is unreachable: never for n >= 1). */
int32_t kmsc_o_mst(const int64_t* w, int32_t n, int32_t* edges, int64_t* dist) {
  if (n <= 1) return 0;
  const size_t nc = (size_t)n * (size_t)(n - 1) / 2;
  mst_cand* c = (mst_cand*)malloc(nc * sizeof(mst_cand));
  size_t t = 0;
  for (int32_t i = 0; i < n; i++)
    for (int32_t j = i + 1; j < n; j++) {
      c[t].d = w[(size_t)i * n + i] + w[(size_t)j * n + j] - 2 * w[(size_t)i * n + j];
      c[t].i = i; c[t].j = j; t++;
    }
  qsort(c, nc, sizeof(mst_cand), mst_cand_cmp);
  kmsc_o_dsu* dsu = kmsc_o_dsu_new(n);
  /* adjacency of the tree as an n x n distance table (-1 = no edge): n is small (sets) */
  int64_t* adj = (int64_t*)malloc((size_t)n * n * sizeof(int64_t));
  for (size_t q = 0; q < (size_t)n * n; q++) adj[q] = -1;
  int32_t taken = 0;
  for (size_t q = 0; q < nc && taken < n - 1; q++) {
    if (kmsc_o_dsu_same(dsu, c[q].i, c[q].j)) continue;
    kmsc_o_dsu_unite(dsu, c[q].i, c[q].j);
    adj[(size_t)c[q].i * n + c[q].j] = c[q].d;
    adj[(size_t)c[q].j * n + c[q].i] = c[q].d;
    taken++;
  }
  int32_t* queue = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  char* seen = (char*)calloc((size_t)n, 1);
  int32_t head = 0, tail = 0, ne = 0;
  queue[tail++] = 0; seen[0] = 1;
  while (head < tail) {
    const int32_t p = queue[head++];
    for (int32_t x = 0; x < n; x++) {
      if (adj[(size_t)p * n + x] < 0 || seen[x]) continue;
      seen[x] = 1;
      edges[2 * ne] = p; edges[2 * ne + 1] = x;
      dist[ne] = adj[(size_t)p * n + x];
      ne++;
      queue[tail++] = x;
    }
  }
  free(queue); free(seen); free(adj); free(c);
  kmsc_o_dsu_free(dsu);
  return ne;
}

/* ------------------------------------------------------------------------- */
/* streamvbyte "0124" (lemire/streamvbyte v0.4.1, published format)           */
/* ------------------------------------------------------------------------- */

size_t kmsc_o_svb0124_max_bytes(uint32_t n) {
  return ((size_t)n + 3) / 4 + (size_t)n * sizeof(uint32_t);
}

size_t kmsc_o_svb0124_encode(const uint32_t* in, uint32_t n, uint8_t* out) {
  uint8_t* ctrl = out;
  uint8_t* data = out + ((size_t)n + 3) / 4;
  uint8_t key = 0;
  int shift = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t v = in[i];
    uint8_t code;
    if (v == 0) code = 0;
    else if (v < (1u << 8)) { code = 1; *data++ = (uint8_t)v; }
    else if (v < (1u << 16)) { code = 2; *data++ = (uint8_t)v; *data++ = (uint8_t)(v >> 8); }
    else { code = 3; memcpy(data, &v, 4); data += 4; } /* little-endian hosts only */
    key |= (uint8_t)(code << shift);
    shift += 2;
    if (shift == 8) { *ctrl++ = key; key = 0; shift = 0; }
  }
  if (shift) *ctrl++ = key;
  return (size_t)(data - out);
}

size_t kmsc_o_svb0124_decode(const uint8_t* in, uint32_t* out, uint32_t n) {
  const uint8_t* ctrl = in;
  const uint8_t* data = in + ((size_t)n + 3) / 4;
  for (uint32_t i = 0; i < n; i++) {
    uint8_t code = (uint8_t)((ctrl[i / 4] >> (2 * (i % 4))) & 3);
    uint32_t v = 0;
    switch (code) {
      case 0: break;
      case 1: v = data[0]; data += 1; break;
      case 2: v = (uint32_t)data[0] | ((uint32_t)data[1] << 8); data += 2; break;
      default: memcpy(&v, data, 4); data += 4; break;
    }
    out[i] = v;
  }
  return (size_t)(data - in);
}
