// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Flat "set key = value" parameter files with the declare / default / abort
// semantics of deal.II's ParameterHandler as the reference uses it
// (reference: src/tools/pde_data.cpp:8-23, :40-83, :110-142): every key is
// declared with a default first, parse_input() overrides, an undeclared key
// aborts the parse, '#' starts a comment.
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace orc {

struct PrmError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class ParameterHandler {
 public:
  void declare_entry(const std::string& key, const std::string& def) { values_[key] = def; }

  void parse_input(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw PrmError("cannot open parameter file " + path);
    std::string line;
    int lineno = 0;
    while (std::getline(in, line)) {
      ++lineno;
      auto hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      line = trim(line);
      if (line.empty()) continue;
      if (line.compare(0, 3, "set") != 0 || (line.size() > 3 && line[3] != ' ' && line[3] != '\t'))
        throw PrmError(path + ":" + std::to_string(lineno) + ": expected 'set key = value'");
      auto eq = line.find('=');
      if (eq == std::string::npos) throw PrmError(path + ":" + std::to_string(lineno) + ": missing '='");
      std::string key = trim(line.substr(3, eq - 3));
      std::string val = trim(line.substr(eq + 1));
      if (!values_.count(key))
        throw PrmError(path + ":" + std::to_string(lineno) + ": no entry with name <" + key + "> was declared");
      values_[key] = val;
    }
  }

  const std::string& get(const std::string& key) const {
    auto it = values_.find(key);
    if (it == values_.end()) throw PrmError("undeclared entry " + key);
    return it->second;
  }
  double get_double(const std::string& key) const { return std::stod(get(key)); }
  bool get_bool(const std::string& key) const {
    const std::string& v = get(key);
    return v == "true" || v == "yes" || v == "on";
  }

 private:
  std::map<std::string, std::string> values_;
  static std::string trim(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && (s[b] == ' ' || s[b] == '\t' || s[b] == '\r')) ++b;
    while (e > b && (s[e - 1] == ' ' || s[e - 1] == '\t' || s[e - 1] == '\r')) --e;
    return s.substr(b, e - b);
  }
};

}  // namespace orc
