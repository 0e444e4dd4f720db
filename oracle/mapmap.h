// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Faithful restatement of the reference's string-keyed pull-back cache MapMap
// (reference: src/tools/mapping.cpp:15-57).  The product does not reproduce it (the
// GPU path recomputes F); it is kept here to pin the reference's five MapMap tests
// (tests/test_tools.cpp:45-103) and to show in the CPU baseline what the original pays.
// Key format quirk reproduced: py[0] is emitted twice ("x0,x1;y0,y0,y1", mapping.cpp:52-55),
// numbers are rendered with the default ostream precision (6 significant digits).
#pragma once
#include <array>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

class MapMap {
 public:
  void set(const std::vector<double>& px, const std::vector<double>& py, double det_jac,
           const std::array<double, 3>& kkt) {
    std::string repr = to_string(px, py);
    determinant_map[repr] = det_jac;
    tensor_map[repr] = kkt;
  }
  void get(const std::vector<double>& px, const std::vector<double>& py, double& det_jac,
           std::array<double, 3>& kkt) const {
    std::string repr = to_string(px, py);
    auto it = determinant_map.find(repr);
    if (it == determinant_map.end()) throw std::runtime_error("Coordinates not present");
    det_jac = it->second;
    kkt = tensor_map.at(repr);
  }
  void get_det_jac(const std::vector<double>& px, const std::vector<double>& py, double& det_jac) const {
    std::string repr = to_string(px, py);
    auto it = determinant_map.find(repr);
    if (it == determinant_map.end()) throw std::runtime_error("Coordinates not present");
    det_jac = it->second;
  }
  static std::string to_string(const std::vector<double>& px, const std::vector<double>& py) {
    std::ostringstream s;
    s << px[0];
    for (size_t i = 1; i < px.size(); ++i) s << "," << px[i];
    s << ";" << py[0];
    for (size_t j = 0; j < py.size(); ++j) s << "," << py[j];
    return s.str();
  }
  unsigned long size() const { return tensor_map.size(); }

 private:
  std::map<std::string, std::array<double, 3>> tensor_map;
  std::map<std::string, double> determinant_map;
};

}  // namespace orc
