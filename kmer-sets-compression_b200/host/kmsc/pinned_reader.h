// kmsc/pinned_reader.h -- the file side of the streaming counter: a reader thread freads
// the (possibly piped) file straight into page-locked buffers, cuts them after whole 2-line FASTA records
// (reference lib/core/kmer_counter.h:163-166) and hands them to the consumer, which gives them to the device
// counter: the copy engine reads the buffer directly and the next chunks are read while the device works.
// Same chunk contents, order and error behaviour as ReadRecordChunks (kmsc/io.h).
#ifndef KMSC_HOST_PINNED_READER_H_
#define KMSC_HOST_PINNED_READER_H_
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "kmsc.h"
#include "kmsc/io.h"
#include "kmsc/status.h"

namespace kmsc {

// `announce(data, n)` (optional) is told about the chunk that will be handed to `sink` NEXT, when it is already
// waiting, before the current one is processed: the device counter starts its copy (kmsc_counter_prefetch).
struct NoAnnounce { void operator()(const char*, std::size_t) const {} };

template <typename Sink, typename Announce = NoAnnounce>
inline Status ReadRecordChunksPinned(const std::string& file_name, const std::string& decompressor,
                                     std::size_t chunk_bytes, Sink sink, Announce announce = Announce(), int n_buffers = 3) {
  const std::size_t cap = chunk_bytes + (static_cast<std::size_t>(8) << 20);
  std::vector<char*> bufs;
  for (int i = 0; i < n_buffers; i++) {
    void* p = nullptr;
    if (kmsc_host_alloc_pinned(cap, &p) != KMSC_OK) break;
    bufs.push_back(static_cast<char*>(p));
  }
  auto release = [&] { for (char* p : bufs) kmsc_host_free_pinned(p); };
  if (bufs.size() < 2) {   // no page-locked memory to be had: pageable chunks, still overlapped
    release();
    return ReadRecordChunksOverlapped(file_name, decompressor, chunk_bytes, sink);
  }
  std::mutex mu;
  std::condition_variable cv;
  std::deque<int> free_q;
  std::deque<std::pair<int, std::size_t>> full_q;
  for (int i = 0; i < static_cast<int>(bufs.size()); i++) free_q.push_back(i);
  bool done = false, stop = false;
  Status reader_status = OkStatus();

  std::thread reader([&] {
    Status st = OkStatus();
    std::FILE* f = nullptr;
    const bool piped = !decompressor.empty();
    f = piped ? popen((decompressor + " < " + file_name).c_str(), "r") : std::fopen(file_name.c_str(), "rb");
    if (f == nullptr) st = InternalError(piped ? "failed to open a sub-process" : "failed to open file");
    auto take_free = [&]() -> int {
      std::unique_lock<std::mutex> l(mu);
      cv.wait(l, [&] { return !free_q.empty() || stop; });
      if (stop) return -1;
      const int b = free_q.front();
      free_q.pop_front();
      return b;
    };
    auto give_full = [&](int b, std::size_t n) {
      std::lock_guard<std::mutex> l(mu);
      full_q.emplace_back(b, n);
      cv.notify_all();
    };
    if (st.ok()) {
      int cur = take_free();
      std::size_t fill = 0, lines = 0, scanned = 0;
      while (cur >= 0) {
        const std::size_t want = std::min<std::size_t>(static_cast<std::size_t>(1) << 23, cap - fill);
        const std::size_t n = want ? std::fread(bufs[cur] + fill, 1, want, f) : 0;
        fill += n;
        if (fill < chunk_bytes && n > 0) continue;
        lines += static_cast<std::size_t>(std::count(bufs[cur] + scanned, bufs[cur] + fill, '\n'));
        scanned = fill;
        if (n == 0 && want > 0) {            // end of the file: whatever is left is the last chunk
          if (fill > 0) give_full(cur, fill);
          break;
        }
        std::size_t cut = 0;                 // bytes that end with an even number of lines
        if (lines >= 2) {
          std::size_t pos = fill;
          int need = lines % 2 == 0 ? 1 : 2;
          while (pos > 0 && need > 0)
            if (bufs[cur][--pos] == '\n') need--;
          cut = pos + 1;
        }
        if (cut == 0) {
          if (fill == cap) { st = InternalError("a FASTA record is longer than the streaming buffer"); break; }
          continue;
        }
        const int nxt = take_free();
        if (nxt < 0) break;
        std::memcpy(bufs[nxt], bufs[cur] + cut, fill - cut);
        give_full(cur, cut);
        cur = nxt;
        fill -= cut;
        lines = 0;
        scanned = 0;
      }
    }
    if (f != nullptr) {
      if (piped) {
        const int exit_status = pclose(f);
        if (st.ok() && exit_status != 0 && !stop)
          st = InternalError("process failed with non-zero exit code: " + std::to_string(exit_status));
      } else {
        std::fclose(f);
      }
    }
    std::lock_guard<std::mutex> l(mu);
    reader_status = st;
    done = true;
    cv.notify_all();
  });

  Status st = OkStatus();
  for (;;) {
    std::pair<int, std::size_t> item, next{-1, 0};
    {
      std::unique_lock<std::mutex> l(mu);
      cv.wait(l, [&] { return !full_q.empty() || done; });
      if (full_q.empty()) break;
      item = full_q.front();
      full_q.pop_front();
      if (!full_q.empty()) next = full_q.front();   // stays in the queue: only this thread takes chunks out
    }
    if (next.first >= 0) announce(bufs[next.first], next.second);
    st = sink(bufs[item.first], item.second);
    std::lock_guard<std::mutex> l(mu);
    if (!st.ok()) { stop = true; cv.notify_all(); break; }
    free_q.push_back(item.first);
    cv.notify_all();
  }
  reader.join();
  release();
  if (!st.ok()) return st;
  return reader_status;
}

}  // namespace kmsc
#endif
