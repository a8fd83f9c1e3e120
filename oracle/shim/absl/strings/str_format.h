// TEST INFRASTRUCTURE ONLY (oracle/). printf-style absl::StrFormat for the
// %s / %d call sites in the reference (std::string and std::atomic<int> args).
#ifndef KMSC_ORACLE_SHIM_STR_FORMAT_H_
#define KMSC_ORACLE_SHIM_STR_FORMAT_H_
#include <atomic>
#include <cstdio>
#include <string>
#include <type_traits>
namespace kmsc_shim {
inline const char* FmtArg(const std::string& s) { return s.c_str(); }
inline const char* FmtArg(const char* s) { return s; }
template <typename T>
inline T FmtArg(const std::atomic<T>& a) { return a.load(); }
template <typename T, typename = std::enable_if_t<std::is_arithmetic<T>::value>>
inline T FmtArg(T v) { return v; }
}  // namespace kmsc_shim
namespace absl {
template <typename... A>
std::string StrFormat(const char* fmt, const A&... a) {
  int n = std::snprintf(nullptr, 0, fmt, kmsc_shim::FmtArg(a)...);
  std::string s(n > 0 ? n : 0, '\0');
  if (n > 0) std::snprintf(&s[0], n + 1, fmt, kmsc_shim::FmtArg(a)...);
  return s;
}
}  // namespace absl
#endif
