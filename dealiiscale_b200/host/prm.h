// Flat "set key = value" parameter files with deal.II ParameterHandler semantics, and the three
// data classes that turn them into named expressions — the C++ host twin of dealiiscale_b200/prm.py.
//
// Mirrors reference src/tools/pde_data.{h,cpp}: MultiscaleData (:7-36), BioData (:38-108),
// TwoPressureData (:110-171).  Every key is declared with its default first; the file overrides;
// an undeclared key aborts the parse; '#' starts a comment.
#pragma once
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ms {

struct ParameterError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class ParameterHandler {
 public:
  void declare_entry(const std::string& key, const std::string& default_value) { entries_[key] = default_value; }
  void parse_input(const std::string& file) {
    std::ifstream in(file);
    if (!in) throw ParameterError("cannot open parameter file <" + file + ">");
    std::string line;
    for (int no = 1; std::getline(in, line); ++no) {
      const auto hash = line.find('#');
      if (hash != std::string::npos) line.resize(hash);
      line = strip(line);
      if (line.empty()) continue;
      const std::string where = file + ":" + std::to_string(no) + ": ";
      if (line.size() < 4 || line.compare(0, 3, "set") != 0 || (line[3] != ' ' && line[3] != '\t'))
        throw ParameterError(where + "expected 'set <name> = <value>'");
      const auto eq = line.find('=');
      if (eq == std::string::npos) throw ParameterError(where + "missing '='");
      const std::string key = strip(line.substr(3, eq - 3)), value = strip(line.substr(eq + 1));
      auto it = entries_.find(key);
      if (it == entries_.end()) throw ParameterError(where + "No entry with name <" + key + "> was declared");
      it->second = value;
    }
  }
  const std::string& get(const std::string& key) const {
    auto it = entries_.find(key);
    if (it == entries_.end()) throw ParameterError("entry <" + key + "> was never declared");
    return it->second;
  }
  double get_double(const std::string& key) const { return std::stod(get(key)); }
  bool get_bool(const std::string& key) const {
    const std::string& v = get(key);
    return v == "true" || v == "yes" || v == "on";
  }

 private:
  std::map<std::string, std::string> entries_;
  static std::string strip(const std::string& s) {
    const char* ws = " \t\r\n";
    const auto b = s.find_first_not_of(ws);
    if (b == std::string::npos) return "";
    return s.substr(b, s.find_last_not_of(ws) - b + 1);
  }
};

using Constants = std::map<std::string, double>;

// MultiscaleData<2>
struct EllipticData {
  ParameterHandler params;
  Constants constants;  // none
  explicit EllipticData(const std::string& file) {
    params.declare_entry("macro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("macro_rhs", " x0^2*sin(x0*x1) + x1^2*sin(x0*x1) + 2*cos(x0 + x1)");
    params.declare_entry("macro_solution", "sin(x0*x1) + cos(x0 + x1)");
    params.declare_entry("macro_bc", "sin(x0*x1) + cos(x0 + x1)");
    params.declare_entry("micro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("micro_rhs", "-sin(x0*x1) - cos(x0 + x1)");
    params.declare_entry("micro_solution", "y0*y1 + exp(x0^2 + x1^2)");
    params.declare_entry("micro_bc", "y0*y1 + exp(x0^2 + x1^2)");
    params.declare_entry("mapping", "(3+x0+x1)*(1.6*y0 - 1.6*y1);(3+x0+x1)*(2.4*y0 + 2.4*y1);");
    params.declare_entry("jac_mapping", "1.6*(3+x0+x1);-1.6*(3+x0+x1);2.4*(3+x0+x1);2.4*(3+x0+x1);");
    params.parse_input(file);
  }
};

// BioData<2>
struct BioData {
  ParameterHandler params;
  Constants constants;
  explicit BioData(const std::string& file) {
    const char* macro_def = "sin(x0*x1) + cos(x0 + x1)";
    const char* macro_rhs_def = " x0^2*sin(x0*x1) + x1^2*sin(x0*x1) + 2*cos(x0 + x1)";
    const char* micro_def = "y0*y1 + exp(x0^2 + x1^2)";
    params.declare_entry("macro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("has_solution", "true");
    for (const char* k : {"solution_u", "bc_u_1", "bc_u_2", "solution_w", "bc_w_1", "bc_w_2"})
      params.declare_entry(k, macro_def);
    params.declare_entry("bulk_rhs_u", macro_rhs_def);
    params.declare_entry("bulk_rhs_w", macro_rhs_def);
    params.declare_entry("inflow_measure", "x0 + x1");
    params.declare_entry("outflow_measure", "x0 + x1");
    params.declare_entry("micro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("solution_v", "-D_1 * sin(x0*x1) - cos(x0 + x1)");
    for (const char* k : {"bulk_rhs_v", "bc_v_1", "bc_v_2", "bc_v_3", "bc_v_4"}) params.declare_entry(k, micro_def);
    params.declare_entry("mapping", "(3+x0+x1)*(1.6*y0 - 1.6*y1);(3+x0+x1)*(2.4*y0 + 2.4*y1);");
    params.declare_entry("jac_mapping", "1.6*(3+x0+x1);-1.6*(3+x0+x1);2.4*(3+x0+x1);2.4*(3+x0+x1);");
    params.declare_entry("num_threads", "-1");
    constants = {{"D_1", 2}, {"D_2", 2}, {"kappa_1", 1}, {"kappa_2", 1}, {"kappa_3", 1}, {"kappa_4", 1}};
    for (const auto& kv : constants) params.declare_entry(kv.first, std::to_string(kv.second));
    params.parse_input(file);
    for (auto& kv : constants) kv.second = params.get_double(kv.first);
  }
  std::vector<double> scalars() const {
    return {constants.at("kappa_1"), constants.at("kappa_2"), constants.at("kappa_3"), constants.at("kappa_4"),
            constants.at("D_2")};
  }
};

// TwoPressureData<2>
struct TwoPressureData {
  ParameterHandler params;
  Constants constants;
  explicit TwoPressureData(const std::string& file) {
    params.declare_entry("macro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("macro_rhs",
                         "-A*(2*(2*x0^2 + 1)*exp(t^2 + x0^2 + x1^2) + 2*(2*x1^2 + 1)*exp(t^2 + x0^2 + x1^2)) - "
                         "12*theta*exp(t^2 + x0^2 + x1^2)");
    params.declare_entry("macro_solution", "exp(t^2 + x0^2 + x1^2)");
    params.declare_entry("macro_bc", "exp(t^2 + x0^2 + x1^2)");
    params.declare_entry("micro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("micro_rhs", "-D*(2*(sin(y0)^2 - cos(y0)^2) + 2*(-sin(y1)^2 + cos(y1)^2))");
    params.declare_entry("micro_solution", "sin(y1)^2 + cos(y0)^2 + 2");
    params.declare_entry("micro_bc_neumann", "-2*D*y0*sin(y0)*cos(y0)");
    params.declare_entry("micro_bc_robin",
                         "2*D*y1*sin(y1)*cos(y1) - kappa*(-R*(sin(y1)^2 + cos(y0)^2 + 2) + p_F + exp(t^2 + x0^2 + x1^2))");
    params.declare_entry("init_rho", "sin(y1)^2 + cos(y0)^2 + 2");
    params.declare_entry("nonlinear", "false");
    constants = {{"A", 0.8}, {"D", 1}, {"theta", 0.5}, {"kappa", 1}, {"p_F", 4}, {"R", 2}};
    for (const auto& kv : constants) params.declare_entry(kv.first, std::to_string(kv.second));
    params.parse_input(file);
    for (auto& kv : constants) kv.second = params.get_double(kv.first);
  }
  std::vector<double> scalars() const {
    return {constants.at("kappa"), constants.at("p_F"), constants.at("R"), constants.at("D")};
  }
};

}  // namespace ms
