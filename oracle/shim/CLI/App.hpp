// TEST INFRASTRUCTURE ONLY (oracle/): the sliver of CLI11 (cli11/2.2.0, conanfile.txt:5)
// that main.cpp:140-161 touches, so the reference's main.cpp compiles unmodified:
//   CLI::App app{"title"}; app.option_defaults()->always_capture_default();
//   app.add_option("-t,--threads", int|double|optional<string>&, "help");
//   app.add_flag("-m,--moving-spheres", bool&, "help");  CLI11_PARSE(app, argc, argv);
// Accepted spellings: -t 4, -t4, --threads 4, --threads=4; flags take no value.
#pragma once
#include <cstdlib>
#include <functional>
#include <iostream>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace CLI {

class ParseError : public std::runtime_error {
 public:
  ParseError(const std::string& m, int code) : std::runtime_error(m), code_(code) {}
  int get_exit_code() const { return code_; }
 private:
  int code_;
};

struct OptionDefaults {
  OptionDefaults* always_capture_default(bool = true) { return this; }
};

class App {
  struct Opt {
    std::vector<std::string> names;
    bool is_flag;
    std::function<void(const std::string&)> set;
    std::string help;
  };
  std::string title_;
  std::vector<Opt> opts_;
  OptionDefaults defaults_;

  static std::vector<std::string> split(const std::string& s) {
    std::vector<std::string> out; std::stringstream ss(s); std::string it;
    while (std::getline(ss, it, ',')) out.push_back(it);
    return out;
  }
  template <typename T> static void assign(T& v, const std::string& s) {
    std::istringstream is(s); T tmp{};
    if (!(is >> tmp) || !is.eof()) throw ParseError("Could not convert: " + s, 104);
    v = tmp;
  }
  static void assign(std::string& v, const std::string& s) { v = s; }
  template <typename T> static void assign(std::optional<T>& v, const std::string& s) { T t{}; assign(t, s); v = t; }

 public:
  explicit App(std::string title = "") : title_(std::move(title)) {}
  OptionDefaults* option_defaults() { return &defaults_; }

  template <typename T> void add_option(const std::string& names, T& var, const std::string& help = "") {
    opts_.push_back({split(names), false, [&var](const std::string& s) { assign(var, s); }, help});
  }
  void add_flag(const std::string& names, bool& var, const std::string& help = "") {
    opts_.push_back({split(names), true, [&var](const std::string&) { var = true; }, help});
  }

  void parse(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      if (a == "-h" || a == "--help") {
        std::ostringstream os; os << title_ << "\nOptions:\n";
        for (auto& o : opts_) { os << "  "; for (auto& n : o.names) os << n << ' '; os << "  " << o.help << "\n"; }
        throw ParseError(os.str(), 0);
      }
      std::string name = a, val; bool has_val = false;
      if (a.rfind("--", 0) == 0) {
        auto eq = a.find('=');
        if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); has_val = true; }
      } else if (a.size() > 2 && a[0] == '-') {
        name = a.substr(0, 2); val = a.substr(2); has_val = true;
      }
      Opt* found = nullptr;
      for (auto& o : opts_) for (auto& n : o.names) if (n == name) found = &o;
      if (!found) throw ParseError("The following argument was not expected: " + a, 109);
      if (found->is_flag) { found->set(""); continue; }
      if (!has_val) {
        if (i + 1 >= argc) throw ParseError(name + ": 1 required", 114);
        val = argv[++i];
      }
      found->set(val);
    }
  }
  int exit(const ParseError& e) const {
    (e.get_exit_code() == 0 ? std::cout : std::cerr) << e.what() << "\n";
    return e.get_exit_code();
  }
};
}  // namespace CLI

#define CLI11_PARSE(app, argc, argv)            \
  try { (app).parse((argc), (argv)); }          \
  catch (const CLI::ParseError& e) { return (app).exit(e); }
