// kmsc/io.h -- line IO with the reference's contract (lib/core/io.h:20-126): plain
// files, or "<decompressor> < file" / "<compressor> > file" through popen; same
// error texts. Unlike the reference the file is read in large blocks.
#ifndef KMSC_HOST_IO_H_
#define KMSC_HOST_IO_H_
#include <cstdio>
#include <string>
#include <vector>

#include "kmsc/status.h"

namespace kmsc {

namespace internal {
inline StatusOr<std::string> ReadAll(const std::string& file_name, const std::string& decompressor) {
  std::FILE* f = nullptr;
  const bool piped = !decompressor.empty();
  if (piped) {
    f = popen((decompressor + " < " + file_name).c_str(), "r");
    if (f == nullptr) return InternalError("failed to open a sub-process");
  } else {
    f = std::fopen(file_name.c_str(), "rb");
    if (f == nullptr) return InternalError("failed to open file");
  }
  std::string s;
  std::vector<char> buf(1 << 20);
  for (;;) {
    const std::size_t n = std::fread(buf.data(), 1, buf.size(), f);
    if (n == 0) break;
    s.append(buf.data(), n);
  }
  if (piped) {
    const int exit_status = pclose(f);
    if (exit_status != 0)
      return InternalError("process failed with non-zero exit code: " + std::to_string(exit_status));
  } else {
    std::fclose(f);
  }
  return s;
}
}  // namespace internal

// Streams a (possibly piped) file in chunks cut after an EVEN number of lines, i.e. whole
// 2-line FASTA records (reference lib/core/kmer_counter.h:163-166), without ever holding the
// whole file: `sink(const char* data, size_t n)` is called once per chunk. A file with an odd
// number of lines ends with a chunk holding an odd number, which the counter rejects with the
// reference's message.
template <typename Sink>
inline Status ReadRecordChunks(const std::string& file_name, const std::string& decompressor, std::size_t chunk_bytes,
                               Sink sink) {
  std::FILE* f = nullptr;
  const bool piped = !decompressor.empty();
  if (piped) {
    f = popen((decompressor + " < " + file_name).c_str(), "r");
    if (f == nullptr) return InternalError("failed to open a sub-process");
  } else {
    f = std::fopen(file_name.c_str(), "rb");
    if (f == nullptr) return InternalError("failed to open file");
  }
  std::string buf;
  buf.reserve(chunk_bytes + (1 << 20));
  std::vector<char> io(1 << 22);
  Status st = OkStatus();
  std::size_t lines = 0, scanned = 0, last_even_end = 0;  // state of the newline scan over buf
  for (;;) {
    const std::size_t n = std::fread(io.data(), 1, io.size(), f);
    if (n > 0) buf.append(io.data(), n);
    if (buf.size() >= chunk_bytes || n == 0) {
      for (; scanned < buf.size(); scanned++)
        if (buf[scanned] == '\n' && (++lines % 2) == 0) last_even_end = scanned + 1;
      if (n == 0) {
        if (!buf.empty()) st = sink(buf.data(), buf.size());
        break;
      }
      if (last_even_end > 0) {
        st = sink(buf.data(), last_even_end);
        if (!st.ok()) break;
        buf.erase(0, last_even_end);
        lines = 0; scanned = 0; last_even_end = 0;
      }
    }
  }
  if (piped) {
    const int exit_status = pclose(f);
    if (st.ok() && exit_status != 0)
      return InternalError("process failed with non-zero exit code: " + std::to_string(exit_status));
  } else {
    std::fclose(f);
  }
  return st;
}

// std::getline semantics: a trailing '\n' does not open another line.
inline std::vector<std::string> SplitLines(const std::string& s) {
  std::vector<std::string> lines;
  std::size_t b = 0;
  while (b < s.size()) {
    std::size_t e = s.find('\n', b);
    if (e == std::string::npos) e = s.size();
    lines.emplace_back(s, b, e - b);
    b = e + 1;
  }
  return lines;
}

inline StatusOr<std::vector<std::string>> ReadLines(const std::string& file_name, const std::string& decompressor) {
  StatusOr<std::string> all = internal::ReadAll(file_name, decompressor);
  if (!all.ok()) return all.status();
  return SplitLines(all.value());
}

inline Status WriteLines(const std::string& file_name, const std::string& compressor,
                         const std::vector<std::string>& lines) {
  std::FILE* f = nullptr;
  const bool piped = !compressor.empty();
  if (piped) {
    f = popen((compressor + " > " + file_name).c_str(), "w");
    if (f == nullptr) return InternalError("failed to open a sub-process");
  } else {
    f = std::fopen(file_name.c_str(), "wb");
    if (f == nullptr) return InternalError("failed to open file");
  }
  for (const std::string& line : lines) {
    if (std::fwrite(line.data(), 1, line.size(), f) != line.size() || std::fputc('\n', f) == EOF) {
      if (piped) pclose(f); else std::fclose(f);
      return InternalError(piped ? "failed to write to the process" : "failed to write to the file");
    }
  }
  if (piped) {
    const int exit_status = pclose(f);
    if (exit_status != 0)
      return InternalError("process failed with non-zero exit code: " + std::to_string(exit_status));
  } else {
    std::fclose(f);
  }
  return OkStatus();
}

}  // namespace kmsc
#endif
