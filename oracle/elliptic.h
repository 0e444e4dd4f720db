// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the reference's elliptic-elliptic two-scale driver
// (solve_elliptic).  Each function cites the reference lines it follows.
// The string-keyed MapMap cache (src/tools/mapping.cpp) is restated as the dense
// [k][cell][q] table it semantically is.
#pragma once
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "expr.h"
#include "fem.h"
#include "prm.h"

namespace orc {
namespace elliptic {

// reference: src/tools/pde_data.cpp:7-36 (MultiscaleData<2>)
struct MultiscaleData {
  ParameterHandler params;
  MacroFunction macro_solution, macro_rhs, macro_bc;
  MultiscaleFunction micro_solution, micro_rhs, micro_bc, mapping, map_jac;

  explicit MultiscaleData(const std::string& param_file) {
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
    params.parse_input(param_file);
    std::map<std::string, double> constants;
    const std::string mv = "x0,x1", msv = "x0,x1,y0,y1";
    macro_rhs.initialize(mv, params.get("macro_rhs"), constants);
    macro_bc.initialize(mv, params.get("macro_bc"), constants);
    macro_solution.initialize(mv, params.get("macro_solution"), constants);
    mapping.initialize(msv, params.get("mapping"), constants, nullptr, false, 2);
    micro_rhs.initialize(msv, params.get("micro_rhs"), constants);
    micro_bc.initialize(msv, params.get("micro_bc"), constants, &mapping);
    micro_solution.initialize(msv, params.get("micro_solution"), constants, &mapping);
    map_jac.initialize(msv, params.get("jac_mapping"), constants, nullptr, false, 4);
  }
};

struct MicroSolver;

// reference: src/elliptic-elliptic/macro.{h,cpp}
struct MacroSolver {
  MultiscaleData& data;
  int refine_level;
  Grid grid;
  SparsityPattern pattern;
  SparseMatrix system_matrix;
  std::vector<double> solution, system_rhs, micro_contribution;
  std::vector<std::array<double, 2>> micro_grid_locations;
  const MicroSolver* micro = nullptr;
  int last_cg_iterations = 0;

  MacroSolver(MultiscaleData& d, int refine) : data(d), refine_level(refine) {}

  // macro.cpp:19-43
  void setup() {
    grid.build(MeshKind::RefineGlobal, 1 << refine_level);
    grid.build_pattern(pattern);
    system_matrix.reinit(pattern);
    solution.assign(grid.n_dofs(), 0.0);
    system_rhs.assign(grid.n_dofs(), 0.0);
    micro_contribution.assign(grid.n_dofs(), 0.0);
    get_dof_locations(micro_grid_locations);
  }
  // macro.cpp:169-175
  void get_dof_locations(std::vector<std::array<double, 2>>& loc) const {
    loc.resize(grid.n_dofs());
    for (int d = 0; d < grid.n_dofs(); ++d) {
      double p[2];
      grid.dof_point(d, p);
      loc[d] = {p[0], p[1]};
    }
  }
  void set_micro_objects(const MicroSolver* m) { micro = m; }
  double get_micro_bulk(int k) const;          // macro.cpp:134-157
  void compute_microscopic_contribution() {    // macro.cpp:162-166
    for (int i = 0; i < grid.n_dofs(); ++i) micro_contribution[i] = get_micro_bulk(i);
  }
  // macro.cpp:57-106
  void assemble_system() {
    const int p = 8;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    compute_microscopic_contribution();
    std::fill(system_matrix.val.begin(), system_matrix.val.end(), 0.0);
    std::fill(system_rhs.begin(), system_rhs.end(), 0.0);
    const double h = grid.h;
    for (int c = 0; c < grid.n_cells(); ++c) {
      double cell_matrix[4][4] = {};
      double cell_rhs[4] = {};
      const auto& cd = grid.cell_dofs[c];
      for (int qj = 0; qj < p; ++qj)
        for (int qi = 0; qi < p; ++qi) {
          double xi = qx[qi], eta = qx[qj];
          double JxW = qw[qi] * qw[qj] * h * h;
          double v[4], g[4][2];
          q1_values(xi, eta, v);
          q1_grads(xi, eta, g);
          double pt[2] = {grid.cell_x0(c) + h * xi, grid.cell_y0(c) + h * eta};
          double micro_cont = 0;
          for (int i = 0; i < 4; ++i) micro_cont += micro_contribution[cd[i]] * v[i];
          double rhs_val = data.macro_rhs.value(pt);
          for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 4; ++j)
              cell_matrix[i][j] += ((g[i][0] / h) * (g[j][0] / h) + (g[i][1] / h) * (g[j][1] / h)) * JxW;
            cell_rhs[i] += v[i] * (micro_cont + rhs_val) * JxW;
          }
        }
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) system_matrix.add(cd[i], cd[j], cell_matrix[i][j]);
        system_rhs[cd[i]] += cell_rhs[i];
      }
    }
    std::map<int, double> bv;
    for (int node = 0; node < grid.n_dofs(); ++node)
      if (grid.node_on_boundary(node)) {
        double pt[2] = {-1.0 + (node % grid.m) * h, -1.0 + (node / grid.m) * h};
        bv[grid.node_dof[node]] = data.macro_bc.value(pt);
      }
    apply_boundary_values(bv, system_matrix, solution, system_rhs);
  }
  // macro.cpp:109-114
  void solve() {
    CGResult r = solve_cg(system_matrix, solution, system_rhs, 1000, 1e-12);
    if (!r.converged) throw std::runtime_error("macro CG: SolverControl::NoConvergence");
    last_cg_iterations = r.iterations;
  }
  void run() {  // macro.cpp:178-181
    assemble_system();
    solve();
  }
  // macro.cpp:117-126; FunctionParser gradients are AutoDerivativeFunction
  // central differences with h = 1e-8.
  void compute_error(double& l2, double& h1) const {
    const int p = 8;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    const double h = grid.h, fd = 1e-8;
    double l2s = 0, h1s = 0;
    for (int c = 0; c < grid.n_cells(); ++c) {
      double cl2 = 0, ch1 = 0;
      const auto& cd = grid.cell_dofs[c];
      for (int qj = 0; qj < p; ++qj)
        for (int qi = 0; qi < p; ++qi) {
          double xi = qx[qi], eta = qx[qj];
          double JxW = qw[qi] * qw[qj] * h * h;
          double v[4], g[4][2];
          q1_values(xi, eta, v);
          q1_grads(xi, eta, g);
          double pt[2] = {grid.cell_x0(c) + h * xi, grid.cell_y0(c) + h * eta};
          double uh = 0, gx = 0, gy = 0;
          for (int i = 0; i < 4; ++i) {
            uh += solution[cd[i]] * v[i];
            gx += solution[cd[i]] * g[i][0] / h;
            gy += solution[cd[i]] * g[i][1] / h;
          }
          double diff = uh - data.macro_solution.value(pt);
          cl2 += diff * diff * JxW;
          double ex[2];
          for (int dd = 0; dd < 2; ++dd) {
            double q1[2] = {pt[0], pt[1]}, q2[2] = {pt[0], pt[1]};
            q1[dd] += fd;
            q2[dd] -= fd;
            ex[dd] = (data.macro_solution.value(q1) - data.macro_solution.value(q2)) / (2 * fd);
          }
          ch1 += ((gx - ex[0]) * (gx - ex[0]) + (gy - ex[1]) * (gy - ex[1])) * JxW;
        }
      l2s += cl2;
      h1s += ch1;
    }
    l2 = std::sqrt(l2s);
    h1 = std::sqrt(h1s);
  }
};

// reference: src/elliptic-elliptic/micro.{h,cpp}
struct MicroSolver {
  MultiscaleData& data;
  int refine_level;
  int fem_q_deg = 12;
  Grid grid;
  SparsityPattern pattern;
  int num_grids = 1;
  std::vector<std::array<double, 2>> grid_locations;
  std::vector<std::vector<double>> solutions, righthandsides;
  std::vector<SparseMatrix> system_matrices;
  const std::vector<double>* macro_solution = nullptr;
  const Grid* macro_grid = nullptr;
  // dense pull-back table (MapMap): [k][cell][q] -> det_jac, kkt00, kkt01, kkt11
  std::vector<double> det_jac, kkt;
  std::vector<double> qx, qw;
  std::vector<int> cg_iterations;
  std::vector<std::vector<double>> cg_history;  // oracle extension: recursive residual per step

  MicroSolver(MultiscaleData& d, int refine) : data(d), refine_level(refine) {}

  void set_grid_locations(const std::vector<std::array<double, 2>>& loc) {  // micro.cpp:242-245
    grid_locations = loc;
    num_grids = static_cast<int>(loc.size());
  }
  void set_macro_solution(const std::vector<double>* sol, const Grid* mg) {  // micro.cpp:97-100
    macro_solution = sol;
    macro_grid = mg;
  }
  // micro.cpp:32-37
  void setup() {
    grid.build(MeshKind::RefineGlobal, 1 << refine_level);  // micro.cpp:40-44
    grid.build_pattern(pattern);                            // micro.cpp:47-53
    const int n = grid.n_dofs();                            // micro.cpp:56-72
    solutions.assign(num_grids, std::vector<double>(n, 0.0));
    righthandsides.assign(num_grids, std::vector<double>(n, 0.0));
    system_matrices.assign(num_grids, SparseMatrix());
    cg_iterations.assign(num_grids, 0);
    cg_history.assign(num_grids, {});
    gauss01(fem_q_deg, qx, qw);
    compute_pullback_objects();
  }
  size_t table_index(int k, int c, int q) const {
    const int nq = fem_q_deg * fem_q_deg;
    return (static_cast<size_t>(k) * grid.n_cells() + c) * nq + q;
  }
  void qpoint(int c, int q, double y[2]) const {
    int qi = q % fem_q_deg, qj = q / fem_q_deg;
    y[0] = grid.cell_x0(c) + grid.h * qx[qi];
    y[1] = grid.cell_y0(c) + grid.h * qx[qj];
  }
  // micro.cpp:75-94
  void compute_pullback_objects() {
    const int nq = fem_q_deg * fem_q_deg;
    det_jac.assign(static_cast<size_t>(num_grids) * grid.n_cells() * nq, 0.0);
    kkt.assign(det_jac.size() * 3, 0.0);
    for (int c = 0; c < grid.n_cells(); ++c)
      for (int k = 0; k < num_grids; ++k)
        for (int q = 0; q < nq; ++q) {
          double y[2], F[4];
          qpoint(c, q, y);
          data.map_jac.mtensor_value(grid_locations[k].data(), y, F);
          double det = F[0] * F[3] - F[1] * F[2];
          // invert(F): deal.II multiplies the adjugate by 1/det
          const double inv_det = 1.0 / det;
          double inv[4] = {F[3] * inv_det, -F[1] * inv_det, -F[2] * inv_det, F[0] * inv_det};
          size_t t = table_index(k, c, q);
          det_jac[t] = det;
          // SymmetricTensor(inv * transpose(inv))
          kkt[3 * t + 0] = inv[0] * inv[0] + inv[1] * inv[1];
          kkt[3 * t + 1] = inv[0] * inv[2] + inv[1] * inv[3];
          kkt[3 * t + 2] = inv[2] * inv[2] + inv[3] * inv[3];
        }
  }
  // micro.cpp:108-136
  void integrate_cell(int c, int k, double cell_matrix[4][4], double cell_rhs[4]) const {
    const int nq = fem_q_deg * fem_q_deg;
    const double h = grid.h;
    for (int q = 0; q < nq; ++q) {
      size_t t = table_index(k, c, q);
      double dj = det_jac[t];
      double k00 = kkt[3 * t], k01 = kkt[3 * t + 1], k11 = kkt[3 * t + 2];
      int qi = q % fem_q_deg, qj = q / fem_q_deg;
      double g[4][2];
      q1_grads(qx[qi], qx[qj], g);
      double JxW = qw[qi] * qw[qj] * h * h;
      for (int i = 0; i < 4; ++i) {
        double gi0 = g[i][0] / h, gi1 = g[i][1] / h;
        // shape_grad(i) * kkt  (Tensor<1> * SymmetricTensor<2>)
        double t0 = gi0 * k00 + gi1 * k01, t1 = gi0 * k01 + gi1 * k11;
        for (int j = 0; j < 4; ++j) cell_matrix[i][j] += (t0 * (g[j][0] / h) + t1 * (g[j][1] / h)) * JxW * dj;
      }
    }
    for (int i = 0; i < 4; ++i)
      for (int q = 0; q < nq; ++q) {
        size_t t = table_index(k, c, q);
        double dj = det_jac[t];
        int qi = q % fem_q_deg, qj = q / fem_q_deg;
        double y[2], mapped[2], v[4];
        qpoint(c, q, y);
        data.mapping.mmap(grid_locations[k].data(), y, mapped);
        double rhs_val = data.micro_rhs.mvalue(grid_locations[k].data(), mapped);
        q1_values(qx[qi], qx[qj], v);
        double JxW = qw[qi] * qw[qj] * h * h;
        cell_rhs[i] += ((*macro_solution)[k] + rhs_val) * v[i] * JxW * dj;
      }
  }
  // micro.cpp:140-171
  void assemble_system() {
    for (int k = 0; k < num_grids; ++k) assemble_one(k);
  }
  void assemble_one(int k) {
    std::fill(righthandsides[k].begin(), righthandsides[k].end(), 0.0);
    std::fill(solutions[k].begin(), solutions[k].end(), 0.0);
    system_matrices[k].reinit(pattern);
    for (int c = 0; c < grid.n_cells(); ++c) {
      double cell_matrix[4][4] = {};
      double cell_rhs[4] = {};
      integrate_cell(c, k, cell_matrix, cell_rhs);
      const auto& cd = grid.cell_dofs[c];
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) system_matrices[k].add(cd[i], cd[j], cell_matrix[i][j]);
        righthandsides[k][cd[i]] += cell_rhs[i];
      }
    }
    std::map<int, double> bv;
    for (int node = 0; node < grid.n_dofs(); ++node)
      if (grid.node_on_boundary(node)) {
        double y[2] = {-1.0 + (node % grid.m) * grid.h, -1.0 + (node / grid.m) * grid.h};
        bv[grid.node_dof[node]] = data.micro_bc.mvalue(grid_locations[k].data(), y);
      }
    apply_boundary_values(bv, system_matrices[k], solutions[k], righthandsides[k]);
  }
  // micro.cpp:175-182
  void solve() {
    for (int k = 0; k < num_grids; ++k) solve_one(k);
  }
  void solve_one(int k) {
    CGResult r = solve_cg(system_matrices[k], solutions[k], righthandsides[k], 1000, 1e-12, &cg_history[k]);
    if (!r.converged) throw std::runtime_error("micro CG: SolverControl::NoConvergence");
    cg_iterations[k] = r.iterations;
  }
  // micro.cpp:224-234: every solution vector becomes the L2 projection of the exact micro solution
  void set_exact_solution() {
    for (int k = 0; k < num_grids; ++k) {
      const double* x = grid_locations[k].data();
      solutions[k] = l2_project(grid, pattern, fem_q_deg, [&](const double* y) { return data.micro_solution.mvalue(x, y); });
    }
  }
  void run() {  // micro.cpp:218-221
    assemble_system();
    solve();
  }
  // Per-grid errors of micro.cpp:189-205 (reference square, no J; H1 with
  // AutoDerivativeFunction central differences of the composed function, h=1e-8).
  void grid_errors(int k, double& l2, double& h1) const {
    const int p = fem_q_deg;
    const double h = grid.h, fd = 1e-8;
    const double* x = grid_locations[k].data();
    double l2s = 0, h1s = 0;
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      double cl2 = 0, ch1 = 0;
      for (int q = 0; q < p * p; ++q) {
        int qi = q % p, qj = q / p;
        double v[4], g[4][2], y[2];
        q1_values(qx[qi], qx[qj], v);
        q1_grads(qx[qi], qx[qj], g);
        qpoint(c, q, y);
        double JxW = qw[qi] * qw[qj] * h * h;
        double uh = 0, gx = 0, gy = 0;
        for (int i = 0; i < 4; ++i) {
          uh += solutions[k][cd[i]] * v[i];
          gx += solutions[k][cd[i]] * g[i][0] / h;
          gy += solutions[k][cd[i]] * g[i][1] / h;
        }
        double diff = uh - data.micro_solution.mvalue(x, y);
        cl2 += diff * diff * JxW;
        double ex[2];
        for (int dd = 0; dd < 2; ++dd) {
          double q1[2] = {y[0], y[1]}, q2[2] = {y[0], y[1]};
          q1[dd] += fd;
          q2[dd] -= fd;
          ex[dd] = (data.micro_solution.mvalue(x, q1) - data.micro_solution.mvalue(x, q2)) / (2 * fd);
        }
        ch1 += ((gx - ex[0]) * (gx - ex[0]) + (gy - ex[1]) * (gy - ex[1])) * JxW;
      }
      l2s += cl2;
      h1s += ch1;
    }
    l2 = std::sqrt(l2s);
    h1 = std::sqrt(h1s);
  }
  // micro.cpp:186-215
  void compute_all_errors(double& l2_error, double& h1_error) const {
    std::vector<double> el2(num_grids), eh1(num_grids);
    for (int k = 0; k < num_grids; ++k) grid_errors(k, el2[k], eh1[k]);
    l2_error = q1_l2_norm(*macro_grid, el2, fem_q_deg);
    h1_error = q1_l2_norm(*macro_grid, eh1, fem_q_deg);
  }
};

// macro.cpp:134-157
inline double MacroSolver::get_micro_bulk(int k) const {
  const MicroSolver& ms = *micro;
  const int p = ms.fem_q_deg;
  const double h = ms.grid.h;
  double integral = 0;
  for (int c = 0; c < ms.grid.n_cells(); ++c)
    for (int q = 0; q < p * p; ++q) {
      int qi = q % p, qj = q / p;
      double v = q1_interp(ms.grid, ms.solutions[k], c, ms.qx[qi], ms.qx[qj]);
      integral += v * ms.det_jac[ms.table_index(k, c, q)] * (ms.qw[qi] * ms.qw[qj] * h * h);
    }
  return integral;
}

struct CycleRecord {
  int cycle;
  double micro_l2, micro_h1, macro_l2, macro_h1;
  std::vector<double> macro_solution;                 // reference dof order
  std::vector<double> micro_contribution;             // coupling vector used by this cycle's macro solve
  std::vector<std::vector<double>> micro_solutions;   // [k][dof], reference dof order
  std::vector<int> micro_cg_iterations;
  int macro_cg_iterations;
};

// reference: src/elliptic-elliptic/manager.{h,cpp}
struct Manager {
  MultiscaleData data;
  MacroSolver macro_solver;
  MicroSolver micro_solver;
  double eps = 1e-4;
  double max_iterations = 1e4;
  int cycle = 0;
  int cycle_limit = -1;   // oracle extension: stop after this many cycles (K-cycle parity), -1 = reference behaviour
  bool keep_solutions = true;
  std::vector<CycleRecord> history;

  Manager(int macro_refinement, int micro_refinement, const std::string& data_file)
      : data(data_file), macro_solver(data, macro_refinement), micro_solver(data, micro_refinement) {}

  void setup() {  // manager.cpp:17-35
    macro_solver.setup();
    std::vector<std::array<double, 2>> loc;
    macro_solver.get_dof_locations(loc);
    micro_solver.set_grid_locations(loc);
    micro_solver.setup();
    micro_solver.set_macro_solution(&macro_solver.solution, &macro_solver.grid);
    macro_solver.set_micro_objects(&micro_solver);
    cycle = 0;
  }
  void fixed_point_iterate() {  // manager.cpp:53-58
    macro_solver.run();
    micro_solver.run();
  }
  void compute_errors(double& old_residual, double& residual) {  // manager.cpp:60-77
    CycleRecord rec;
    rec.cycle = cycle;
    macro_solver.compute_error(rec.macro_l2, rec.macro_h1);
    micro_solver.compute_all_errors(rec.micro_l2, rec.micro_h1);
    rec.micro_contribution = macro_solver.micro_contribution;
    rec.macro_cg_iterations = macro_solver.last_cg_iterations;
    rec.micro_cg_iterations = micro_solver.cg_iterations;
    if (keep_solutions) {
      rec.macro_solution = macro_solver.solution;
      rec.micro_solutions = micro_solver.solutions;
    }
    old_residual = residual;
    residual = rec.micro_l2 + rec.macro_l2;
    history.push_back(std::move(rec));
  }
  void run() {  // manager.cpp:37-51
    double old_residual = 1, residual = 0;
    while (std::fabs(old_residual - residual) > eps) {
      fixed_point_iterate();
      compute_errors(old_residual, residual);
      cycle++;
      if (cycle > max_iterations) break;
      if (cycle_limit >= 0 && cycle >= cycle_limit) break;
    }
  }
  // ConvergenceTable::write_text as configured at manager.cpp:79-88.
  std::string convergence_table() const {
    std::string out = "cycle cells dofs    mL2       mH1       ML2       MH1    \n";
    char buf[256];
    for (const auto& r : history) {
      std::snprintf(buf, sizeof buf, "%5d %5d %4d %.3e %.3e %.3e %.3e\n", r.cycle, macro_solver.grid.n_cells(),
                    macro_solver.grid.n_dofs(), r.micro_l2, r.micro_h1, r.macro_l2, r.macro_h1);
      out += buf;
    }
    return out;
  }
};

}  // namespace elliptic
}  // namespace orc
