#include "problem_programs.h"

#include <map>

namespace msgpu {

namespace {
// keep in sync with msgpu_function_slot (include/msgpu.h)
enum { FN_MAPPING = 0, FN_JACOBIAN, FN_RHS, FN_BC, FN_SOLUTION, FN_BC_V1, FN_BC_V2, FN_BC_V3, FN_BC_V4, FN_INIT };

std::map<std::string, double> const_map(const FnDef& f) {
  std::map<std::string, double> m;
  for (size_t i = 0; i < f.const_names.size(); ++i) m[f.const_names[i]] = f.const_values[i];
  return m;
}
}  // namespace

int build_problem_programs(int problem, const FnDef* fn, ProblemPrograms& out, std::string& err) {
  out = ProblemPrograms();
  ProgramSet& ps = out.ps;
  auto need = [&](int slot, const char* what) {
    if (fn[slot].set) return true;
    err = std::string("function not set: ") + what;
    return false;
  };
  try {
    if (problem == 0 || problem == 1) {
      if (!need(FN_MAPPING, "mapping") || !need(FN_JACOBIAN, "jac_mapping") || !need(FN_RHS, "micro rhs")) return 1;
      out.has_map = 1;
      const FnDef& fm = fn[FN_MAPPING];
      const FnDef& fj = fn[FN_JACOBIAN];
      std::vector<int> map = ps.parse_function(fm.expr, const_map(fm), fm.time_dep, 2);
      std::vector<int> jac = ps.parse_function(fj.expr, const_map(fj), fj.time_dep, 4);
      auto mapped = [&](int slot) {
        const FnDef& f = fn[slot];
        return ps.parse_function(f.expr, const_map(f), f.time_dep, 1, map[0], map[1])[0];
      };
      std::vector<int> vol = jac;
      const int g_node = mapped(FN_RHS);
      vol.push_back(g_node);
      out.prog_vol = ps.add_point_program(vol, MODE_LINES);
      out.prog_vol_F = ps.add_point_program(jac, MODE_LINES);
      out.prog_vol_G = ps.add_point_program({g_node}, MODE_LINES);
      if (problem == 0) {
        if (!need(FN_BC, "micro_bc")) return 1;
        out.prog_bv = ps.add_point_program({mapped(FN_BC)}, MODE_LINES);
      } else {
        // reference boundary ids (bio micro.cpp:76-93): x=-1 INFLOW bc_v_1, x=+1 OUTFLOW bc_v_2,
        // y=+1 TOP bc_v_3, y=-1 BOTTOM bc_v_4
        const int side_fn[4] = {FN_BC_V1, FN_BC_V2, FN_BC_V4, FN_BC_V3};
        for (int s = 0; s < 4; ++s) {
          if (!need(side_fn[s], "bc_v_*")) return 1;
          std::vector<int> o = jac;
          o.push_back(mapped(side_fn[s]));
          out.prog_face[s] = ps.add_point_program(o, MODE_LINES);
        }
      }
      if (fn[FN_SOLUTION].set) out.prog_err = ps.add_point_program({mapped(FN_SOLUTION)}, MODE_POINTS);
      out.jac_depends_on_x = 0;
      out.jac_depends_on_y = 0;
      for (int j : jac) {
        if (ps.node_dep(j) & DEP_X) out.jac_depends_on_x = 1;
        if (ps.node_dep(j) & (DEP_Y0 | DEP_Y1)) out.jac_depends_on_y = 1;
      }
    } else {
      if (!need(FN_RHS, "micro_rhs") || !need(FN_BC_V1, "micro_bc_robin") || !need(FN_BC_V2, "micro_bc_neumann"))
        return 1;
      out.has_map = 0;
      out.jac_depends_on_x = 0;
      out.jac_depends_on_y = 0;
      auto plain = [&](int slot) {
        const FnDef& f = fn[slot];
        return ps.parse_function(f.expr, const_map(f), f.time_dep, 1)[0];
      };
      out.prog_vol = ps.add_point_program({plain(FN_RHS)}, MODE_LINES);
      out.prog_vol_G = out.prog_vol;
      // reference boundary ids (rho_solver.cpp:40-50): |y1| = 1 NEUMANN, the rest (|y0| = 1) ROBIN
      const int robin = plain(FN_BC_V1), neumann = plain(FN_BC_V2);
      out.prog_face[0] = ps.add_point_program({robin}, MODE_LINES);
      out.prog_face[1] = out.prog_face[0];
      out.prog_face[2] = ps.add_point_program({neumann}, MODE_LINES);
      out.prog_face[3] = out.prog_face[2];
      if (fn[FN_INIT].set) out.prog_init = ps.add_point_program({plain(FN_INIT)}, MODE_POINTS);
      if (fn[FN_SOLUTION].set) out.prog_err = ps.add_point_program({plain(FN_SOLUTION)}, MODE_POINTS);
    }
    ps.finalize();
  } catch (const std::exception& e) {
    err = e.what();
    return 2;
  }
  return 0;
}

}  // namespace msgpu
