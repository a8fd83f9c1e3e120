// kmsc/streamvbyte0124.h -- the "0124" variant of streamvbyte as the reference uses it
// for SPSS string lengths (lib/core/kmer_set_compact.h:258-263, 272; third party
// lemire/streamvbyte v0.4.1, not vendored): ceil(n/4) control bytes first, two bits
// per value (first value in the low bits), code 0/1/2/3 = 0/1/2/4 little-endian data
// bytes. The reference never serialises these bytes, only round-trips them in memory.
#ifndef KMSC_HOST_STREAMVBYTE0124_H_
#define KMSC_HOST_STREAMVBYTE0124_H_
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace kmsc {

inline std::size_t Svb0124MaxBytes(std::size_t n) { return (n + 3) / 4 + n * 4; }

inline std::vector<std::uint8_t> Svb0124Encode(const std::vector<std::uint32_t>& in) {
  const std::size_t n = in.size();
  std::vector<std::uint8_t> out(Svb0124MaxBytes(n), 0);
  std::size_t d = (n + 3) / 4;
  for (std::size_t i = 0; i < n; i++) {
    const std::uint32_t v = in[i];
    unsigned code, bytes;
    if (v == 0) { code = 0; bytes = 0; }
    else if (v < (1u << 8)) { code = 1; bytes = 1; }
    else if (v < (1u << 16)) { code = 2; bytes = 2; }
    else { code = 3; bytes = 4; }
    out[i / 4] |= static_cast<std::uint8_t>(code << (2 * (i % 4)));
    for (unsigned b = 0; b < bytes; b++) out[d++] = static_cast<std::uint8_t>(v >> (8 * b));
  }
  out.resize(d);
  out.shrink_to_fit();
  return out;
}

inline std::vector<std::uint32_t> Svb0124Decode(const std::vector<std::uint8_t>& in, std::size_t n) {
  std::vector<std::uint32_t> out(n);
  std::size_t d = (n + 3) / 4;
  for (std::size_t i = 0; i < n; i++) {
    const unsigned code = (in[i / 4] >> (2 * (i % 4))) & 3u;
    const unsigned bytes = code == 3 ? 4 : code;
    std::uint32_t v = 0;
    for (unsigned b = 0; b < bytes; b++) v |= static_cast<std::uint32_t>(in[d++]) << (8 * b);
    out[i] = v;
  }
  return out;
}

}  // namespace kmsc
#endif
