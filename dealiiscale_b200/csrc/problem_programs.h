// Which expressions of a .prm file feed which VM program, per driver.
//
// Mirrors how the reference wires its parsed functions together:
//   elliptic  src/tools/pde_data.cpp:25-35   (bc and solution get &micro.mapping as mapper; rhs is mapped
//             by the caller, src/elliptic-elliptic/micro.cpp:129-131)
//   bio       src/tools/pde_data.cpp:84-107  (bulk_rhs_v and bc_v_* are evaluated at mapping.mmap(x,y),
//             src/bio-multiscale/micro.cpp:213-215, :229-230, :241-267)
//   parabolic src/tools/pde_data.cpp:144-158 (no mapping; functions of (x,y,t))
#pragma once
#include <string>
#include <vector>

#include "expr_compile.h"

namespace msgpu {

struct FnDef {
  bool set = false;
  std::string expr;
  std::vector<std::string> const_names;
  std::vector<double> const_values;
  bool time_dep = false;
};

struct ProblemPrograms {
  ProgramSet ps;
  int prog_vol = -1;             // volume quadrature points: F00 F01 F10 F11 G   (no mapping: G)
  int prog_vol_F = -1;           // F00 F01 F10 F11 only (run once per cell when F does not depend on y)
  int prog_vol_G = -1;           // G only
  int prog_bv = -1;              // elliptic Dirichlet data at boundary vertices
  int prog_err = -1;             // exact solution at arbitrary points (MODE_POINTS)
  int prog_init = -1;            // parabolic initial condition at arbitrary points
  int prog_face[4] = {-1, -1, -1, -1};  // sides x=-1, x=+1, y=-1, y=+1: F00..F11 BC   (no mapping: BC)
  int has_map = 0;
  int jac_depends_on_x = 1;
  int jac_depends_on_y = 1;      // 0: K and det(F) are constant per system, stiffness moments are closed-form
};

// problem: 0 elliptic, 1 bio, 2 parabolic (msgpu_problem); fn indexed by msgpu_function_slot.
// Returns 0, or 1 (missing function) / 2 (parse error) with a message in err.
int build_problem_programs(int problem, const FnDef* fn, ProblemPrograms& out, std::string& err);

}  // namespace msgpu
