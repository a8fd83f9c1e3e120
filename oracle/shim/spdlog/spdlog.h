// TEST INFRASTRUCTURE ONLY (oracle/). Minimal spdlog: "{}" formatting to stderr,
// plus a hook the oracle driver uses to time the reference's phases (every
// debug() call reports its format string and text to kmsc_shim::log_hook first; the hook
// may throw to leave a long-running reference routine after a given phase).
#ifndef KMSC_ORACLE_SHIM_SPDLOG_H_
#define KMSC_ORACLE_SHIM_SPDLOG_H_
#include <atomic>
#include <cstdio>
#include <sstream>
#include <string>
namespace kmsc_shim {
using LogHook = void (*)(const char* fmt, const char* formatted);
inline LogHook& log_hook() { static LogHook h = nullptr; return h; }
inline int& log_level() { static int l = 2; return l; }  // 1 debug, 2 info, 4 err
template <typename T> const T& Unwrap(const T& v) { return v; }
template <typename T> T Unwrap(const std::atomic<T>& v) { return v.load(); }
inline void FormatTo(std::ostringstream& os, const char* f) { os << f; }
template <typename A, typename... R>
void FormatTo(std::ostringstream& os, const char* f, const A& a, const R&... r) {
  for (; *f; ++f) {
    if (f[0] == '{' && f[1] == '}') { os << Unwrap(a); FormatTo(os, f + 2, r...); return; }
    os << *f;
  }
}
template <typename... A>
void Log(int level, const char* tag, const char* fmt, const A&... a) {
  if (level < log_level()) return;
  std::ostringstream os;
  FormatTo(os, fmt, a...);
  std::fprintf(stderr, "[%s] %s\n", tag, os.str().c_str());
}
}  // namespace kmsc_shim
namespace spdlog {
namespace level { enum level_enum { trace = 0, debug = 1, info = 2, warn = 3, err = 4 }; }
inline void set_level(level::level_enum l) { kmsc_shim::log_level() = l; }
template <typename... A> void debug(const char* fmt, const A&... a) {
  if (kmsc_shim::log_hook()) {
    std::ostringstream os;
    kmsc_shim::FormatTo(os, fmt, a...);
    kmsc_shim::log_hook()(fmt, os.str().c_str());
  }
  kmsc_shim::Log(1, "debug", fmt, a...);
}
template <typename... A> void info(const char* fmt, const A&... a) { kmsc_shim::Log(2, "info", fmt, a...); }
template <typename... A> void error(const char* fmt, const A&... a) { kmsc_shim::Log(4, "error", fmt, a...); }
}  // namespace spdlog
#endif
