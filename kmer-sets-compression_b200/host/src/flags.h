// flags.h -- tiny command-line parser with the reference executables' conventions
// (Abseil flags: --name=value, --name value, --bool / --nobool / --bool=false,
// everything else positional; reference lib/flags.h, src/*.cc). Logging mirrors the
// reference's spdlog stderr lines ("[info] ...").
#ifndef KMSC_HOST_FLAGS_H_
#define KMSC_HOST_FLAGS_H_
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace kmsc_cli {

struct Flags {
  std::map<std::string, std::string> values;
  std::vector<std::string> positional;

  std::string Str(const std::string& name, const std::string& def) const {
    auto it = values.find(name);
    return it == values.end() ? def : it->second;
  }
  int Int(const std::string& name, int def) const {
    auto it = values.find(name);
    return it == values.end() ? def : std::atoi(it->second.c_str());
  }
  bool Bool(const std::string& name, bool def) const {
    auto it = values.find(name);
    if (it == values.end()) return def;
    return !(it->second == "false" || it->second == "0" || it->second == "no");
  }
};

inline Flags ParseFlags(int argc, char** argv, const std::vector<std::string>& bool_flags) {
  Flags f;
  auto is_bool = [&](const std::string& n) {
    for (const std::string& b : bool_flags) if (b == n) return true;
    return false;
  };
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a.rfind("--", 0) != 0) { f.positional.push_back(a); continue; }
    a = a.substr(2);
    const std::size_t eq = a.find('=');
    if (eq != std::string::npos) { f.values[a.substr(0, eq)] = a.substr(eq + 1); continue; }
    if (is_bool(a)) { f.values[a] = "true"; continue; }
    if (a.rfind("no", 0) == 0 && is_bool(a.substr(2))) { f.values[a.substr(2)] = "false"; continue; }
    if (i + 1 < argc) f.values[a] = argv[++i];
  }
  return f;
}

inline void Info(const std::string& s) { std::fprintf(stderr, "[info] %s\n", s.c_str()); }
inline void Error(const std::string& s) { std::fprintf(stderr, "[error] %s\n", s.c_str()); }

}  // namespace kmsc_cli
#endif
