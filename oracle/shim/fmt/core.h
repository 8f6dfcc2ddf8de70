// TEST INFRASTRUCTURE ONLY (oracle/): the two fmt calls main.cpp makes (fmt/9.0.0 in
// conanfile.txt:6): fmt::format("...{}...", args) at main.cpp:111,130 and
// fmt::print(stderr, "...{}...", args) at main.cpp:134. Only "{}" placeholders are used.
#pragma once
#include <cstdio>
#include <sstream>
#include <string>
#include <string_view>

namespace fmt {
namespace shim_detail {
inline void emit(std::ostringstream& os, std::string_view f) { os << f; }
template <typename A, typename... R>
void emit(std::ostringstream& os, std::string_view f, const A& a, const R&... rest) {
  auto p = f.find("{}");
  if (p == std::string_view::npos) { os << f; return; }
  os << f.substr(0, p) << a;
  emit(os, f.substr(p + 2), rest...);
}
}  // namespace shim_detail

template <typename... Args>
std::string format(std::string_view f, const Args&... args) {
  std::ostringstream os;
  shim_detail::emit(os, f, args...);
  return os.str();
}
template <typename... Args>
void print(std::FILE* out, std::string_view f, const Args&... args) {
  std::string s = format(f, args...);
  std::fwrite(s.data(), 1, s.size(), out);
}
}  // namespace fmt
