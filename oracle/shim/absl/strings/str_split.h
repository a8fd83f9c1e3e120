// TEST INFRASTRUCTURE ONLY (oracle/). absl::StrSplit(text, char) -> vector<string>
// (keeps empty pieces, like Abseil's default).
#ifndef KMSC_ORACLE_SHIM_STR_SPLIT_H_
#define KMSC_ORACLE_SHIM_STR_SPLIT_H_
#include <string>
#include <vector>
namespace absl {
inline std::vector<std::string> StrSplit(const std::string& s, char d) {
  std::vector<std::string> out;
  std::size_t b = 0;
  while (true) {
    std::size_t e = s.find(d, b);
    if (e == std::string::npos) { out.emplace_back(s, b); break; }
    out.emplace_back(s, b, e - b);
    b = e + 1;
  }
  return out;
}
}  // namespace absl
#endif
