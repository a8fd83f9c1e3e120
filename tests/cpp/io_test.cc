// Host-only checks of the streaming file readers (host/kmsc/io.h: ReadRecordChunks, ReadRecordChunksOverlapped):
// chunks are cut after whole 2-line FASTA records (reference lib/core/kmer_counter.h:163-166), their
// concatenation is the file, a file with an odd number of lines ends with an odd chunk, a piped
// decompressor gives the same chunks, a failing sink stops the stream. No device is touched.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <vector>

#include "kmsc/io.h"

using namespace kmsc;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)

static std::size_t count_lines(const std::string& s) { return static_cast<std::size_t>(std::count(s.begin(), s.end(), '\n')); }

template <typename Reader>
static void check_reader(const std::string& path, const std::string& content, std::size_t chunk_bytes, const std::string& decompressor,
                         Reader reader) {
  std::vector<std::string> chunks;
  Status st = reader(path, decompressor, chunk_bytes, [&](const char* d, std::size_t n) -> Status {
    chunks.emplace_back(d, n);
    return OkStatus();
  });
  CHECK(st.ok());
  std::string all;
  for (const std::string& c : chunks) all += c;
  CHECK(all == content);
  if (content.size() > (std::size_t(9) << 20) && chunk_bytes <= (std::size_t(1) << 20)) CHECK(chunks.size() >= 2);
  for (std::size_t i = 0; i + 1 < chunks.size(); i++) {
    CHECK(!chunks[i].empty() && chunks[i].back() == '\n');
    CHECK(count_lines(chunks[i]) % 2 == 0);
    CHECK(chunks[i].size() >= chunk_bytes || chunks.size() == 1);
  }
}

int main() {
  const std::string dir = std::string(std::getenv("TMPDIR") ? std::getenv("TMPDIR") : "/tmp");
  std::mt19937 gen(7);
  for (int variant = 0; variant < 4; variant++) {
    // records with lines of very different lengths, with / without a trailing newline, even / odd line counts
    std::string content;
    const int n_lines = 60000 + variant % 2;   // ~11 MB: several 4 MB read blocks; variant 1, 3: an odd number of lines
    for (int i = 0; i < n_lines; i++) {
      const int len = (i % 2 == 0) ? static_cast<int>(gen() % 20) + 1 : static_cast<int>(gen() % 700);
      content += std::string(static_cast<std::size_t>(len), "ACGT>"[gen() % 5]);
      if (i + 1 < n_lines || variant < 2) content += '\n';
    }
    const std::string path = dir + "/kmsc_io_test_" + std::to_string(variant) + ".txt";
    { std::ofstream f(path, std::ios::binary); f << content; }
    for (std::size_t chunk : {std::size_t(1), std::size_t(100), std::size_t(5000), std::size_t(1) << 20}) {
      auto plain = [](const std::string& p, const std::string& d, std::size_t c, auto sink) { return ReadRecordChunks(p, d, c, sink); };
      auto overlapped = [](const std::string& p, const std::string& d, std::size_t c, auto sink) { return ReadRecordChunksOverlapped(p, d, c, sink); };
      check_reader(path, content, chunk, "", plain);
      check_reader(path, content, chunk, "", overlapped);
      check_reader(path, content, chunk, "cat", plain);
      check_reader(path, content, chunk, "cat", overlapped);
    }
    // a failing sink stops the stream and its status comes back (the file is read in 4 MB blocks: one chunk here)
    int calls = 0;
    Status st = ReadRecordChunksOverlapped(path, "", 100, [&](const char*, std::size_t) -> Status {
      calls++;
      return InternalError("stop here");
    });
    CHECK(!st.ok() && calls == 1);
    std::remove(path.c_str());
  }
  CHECK(!ReadRecordChunks(dir + "/kmsc_io_test_missing.txt", "", 100, [](const char*, std::size_t) { return OkStatus(); }).ok());
  CHECK(!ReadRecordChunksOverlapped(dir + "/kmsc_io_test_missing.txt", "", 100, [](const char*, std::size_t) { return OkStatus(); }).ok());
  CHECK(!ReadRecordChunks("/dev/null", "false", 100, [](const char*, std::size_t) { return OkStatus(); }).ok());
  std::printf(failures ? "io_test: %d failures\n" : "io_test: ok\n", failures);
  return failures ? 1 : 0;
}
