/* TEST INFRASTRUCTURE ONLY (oracle/). Declarations of the three streamvbyte
 * entry points the reference calls (lib/core/kmer_set_compact.h:258-263, 272).
 * lemire/streamvbyte v0.4.1 (extern/install.sh:69) is not vendored under
 * /root/reference and cannot be fetched; the definitions live in
 * oracle/kmsc_oracle.c and restate the published "0124" format:
 * ceil(n/4) control bytes first (2 bits per value, first value in the low
 * bits), then the data bytes; code 0/1/2/3 = 0/1/2/4 little-endian bytes.
 * Byte-level parity with upstream: UNPINNED (bytes are never serialised by
 * the reference; only encode->decode round trips are observable). */
#ifndef KMSC_ORACLE_SHIM_STREAMVBYTE_H_
#define KMSC_ORACLE_SHIM_STREAMVBYTE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t kmsc_o_svb0124_max_bytes(uint32_t n);
size_t kmsc_o_svb0124_encode(const uint32_t* in, uint32_t n, uint8_t* out);
size_t kmsc_o_svb0124_decode(const uint8_t* in, uint32_t* out, uint32_t n);
static inline size_t streamvbyte_max_compressedbytes(uint32_t n) { return kmsc_o_svb0124_max_bytes(n); }
static inline size_t streamvbyte_encode_0124(const uint32_t* in, uint32_t n, uint8_t* out) {
  return kmsc_o_svb0124_encode(in, n, out);
}
static inline size_t streamvbyte_decode_0124(const uint8_t* in, uint32_t* out, uint32_t n) {
  return kmsc_o_svb0124_decode(in, out, n);
}
#ifdef __cplusplus
}
#endif
#endif
