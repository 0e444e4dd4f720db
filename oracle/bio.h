// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the reference's bio-multiscale driver (solve_biomath).  Each function
// cites the reference lines it follows.  cache_mappings is false in the reference
// (src/bio-multiscale/micro.cpp:51), so F, K, J are recomputed from the parsed Jacobian at every
// quadrature point through MicroFEMObjects::get_map_* (src/tools/pde_data.cpp:174-209).
#pragma once
#include <string>
#include <thread>
#include <vector>

#include "expr.h"
#include "fem.h"
#include "prm.h"

namespace orc {
namespace bio {

// reference: src/tools/pde_data.cpp:38-108 (BioData<2>)
struct BioData {
  ParameterHandler params;
  MacroFunction solution_u, bulk_rhs_u, bc_u_1, bc_u_2, inflow_measure, outflow_measure;
  MacroFunction solution_w, bulk_rhs_w, bc_w_1, bc_w_2;
  MultiscaleFunction solution_v, bulk_rhs_v, bc_v_1, bc_v_2, bc_v_3, bc_v_4, mapping, map_jac;
  std::map<std::string, double> constants;

  explicit BioData(const std::string& param_file) {
    const char* macro_def = "sin(x0*x1) + cos(x0 + x1)";
    const char* macro_rhs_def = " x0^2*sin(x0*x1) + x1^2*sin(x0*x1) + 2*cos(x0 + x1)";
    const char* micro_def = "y0*y1 + exp(x0^2 + x1^2)";
    params.declare_entry("macro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("has_solution", "true");
    params.declare_entry("solution_u", macro_def);
    params.declare_entry("bulk_rhs_u", macro_rhs_def);
    params.declare_entry("bc_u_1", macro_def);
    params.declare_entry("bc_u_2", macro_def);
    params.declare_entry("inflow_measure", "x0 + x1");
    params.declare_entry("outflow_measure", "x0 + x1");
    params.declare_entry("solution_w", macro_def);
    params.declare_entry("bulk_rhs_w", macro_rhs_def);
    params.declare_entry("bc_w_1", macro_def);
    params.declare_entry("bc_w_2", macro_def);
    params.declare_entry("micro_geometry", "[-1,1]x[-1,1]");
    params.declare_entry("solution_v", "-D_1 * sin(x0*x1) - cos(x0 + x1)");
    params.declare_entry("bulk_rhs_v", micro_def);
    params.declare_entry("bc_v_1", micro_def);
    params.declare_entry("bc_v_2", micro_def);
    params.declare_entry("bc_v_3", micro_def);
    params.declare_entry("bc_v_4", micro_def);
    params.declare_entry("mapping", "(3+x0+x1)*(1.6*y0 - 1.6*y1);(3+x0+x1)*(2.4*y0 + 2.4*y1);");
    params.declare_entry("jac_mapping", "1.6*(3+x0+x1);-1.6*(3+x0+x1);2.4*(3+x0+x1);2.4*(3+x0+x1);");
    params.declare_entry("num_threads", "-1");
    constants = {{"D_1", 2}, {"D_2", 2}, {"kappa_1", 1}, {"kappa_2", 1}, {"kappa_3", 1}, {"kappa_4", 1}};
    for (const auto& kv : constants) params.declare_entry(kv.first, std::to_string(kv.second));
    params.parse_input(param_file);
    for (auto& kv : constants) kv.second = params.get_double(kv.first);
    const std::string mv = "x0,x1", msv = "x0,x1,y0,y1";
    solution_u.initialize(mv, params.get("solution_u"), constants);
    bulk_rhs_u.initialize(mv, params.get("bulk_rhs_u"), constants);
    bc_u_1.initialize(mv, params.get("bc_u_1"), constants);
    bc_u_2.initialize(mv, params.get("bc_u_2"), constants);
    inflow_measure.initialize(mv, params.get("inflow_measure"), constants);
    outflow_measure.initialize(mv, params.get("outflow_measure"), constants);
    solution_w.initialize(mv, params.get("solution_w"), constants);
    bulk_rhs_w.initialize(mv, params.get("bulk_rhs_w"), constants);
    bc_w_1.initialize(mv, params.get("bc_w_1"), constants);
    bc_w_2.initialize(mv, params.get("bc_w_2"), constants);
    mapping.initialize(msv, params.get("mapping"), constants, nullptr, false, 2);
    map_jac.initialize(msv, params.get("jac_mapping"), constants, nullptr, false, 4);
    solution_v.initialize(msv, params.get("solution_v"), constants, &mapping);
    bulk_rhs_v.initialize(msv, params.get("bulk_rhs_v"), constants);
    bc_v_1.initialize(msv, params.get("bc_v_1"), constants);
    bc_v_2.initialize(msv, params.get("bc_v_2"), constants);
    bc_v_3.initialize(msv, params.get("bc_v_3"), constants);
    bc_v_4.initialize(msv, params.get("bc_v_4"), constants);
  }
  // pde_data.cpp:174-186
  void get_map_data(const double* px, const double* py, double& det_jac, double kkt[3]) const {
    double F[4];
    map_jac.mtensor_value(px, py, F);
    det_jac = F[0] * F[3] - F[2] * F[1];
    const double inv_det = 1.0 / det_jac;
    const double i00 = F[3] * inv_det, i01 = -F[1] * inv_det, i10 = -F[2] * inv_det, i11 = F[0] * inv_det;
    kkt[0] = i00 * i00 + i01 * i01;
    kkt[1] = i00 * i10 + i01 * i11;
    kkt[2] = i10 * i10 + i11 * i11;
  }
  // pde_data.cpp:189-197
  double get_map_det_jac(const double* px, const double* py) const {
    double F[4];
    map_jac.mtensor_value(px, py, F);
    return F[0] * F[3] - F[2] * F[1];
  }
  // pde_data.cpp:200-209: |F * R * n|, R = [[0,-1],[1,0]] (pde_data.h:68-71)
  double get_map_det_jac_bc(const double* px, const double* py, const double* normal) const {
    double F[4];
    map_jac.mtensor_value(px, py, F);
    const double rn0 = -normal[1], rn1 = normal[0];
    const double v0 = F[0] * rn0 + F[1] * rn1, v1 = F[2] * rn0 + F[3] * rn1;
    return std::sqrt(v0 * v0 + v1 * v1);
  }
};

enum { INFLOW_BOUNDARY = 0, OUTFLOW_BOUNDARY = 1, TOP_NEUMANN = 2, BOTTOM_NEUMANN = 3 };

// boundary id of face f (0:x-min 1:x-max 2:y-min 3:y-max) — bio micro.cpp:76-93
inline int micro_face_id(int f) {
  switch (f) {
    case 0: return INFLOW_BOUNDARY;
    case 1: return OUTFLOW_BOUNDARY;
    case 2: return BOTTOM_NEUMANN;
    default: return TOP_NEUMANN;
  }
}
static const double FACE_NORMAL[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
// unit-cell point of face quadrature abscissa s on face f
inline void face_point(int f, double s, double& xi, double& eta) {
  switch (f) {
    case 0: xi = 0; eta = s; break;
    case 1: xi = 1; eta = s; break;
    case 2: xi = s; eta = 0; break;
    default: xi = s; eta = 1; break;
  }
}

// reference: src/bio-multiscale/micro.{h,cpp}
struct MicroSolver {
  BioData& data;
  int n_cells_axis;
  int fem_q_deg = 12;
  Grid grid;
  SparsityPattern pattern;
  int num_grids = 1;
  std::vector<std::array<double, 2>> grid_locations;
  std::vector<std::vector<double>> solutions, old_solutions, righthandsides;
  std::vector<SparseMatrix> system_matrices;
  const std::vector<double>*sol_u = nullptr, *sol_w = nullptr;
  const Grid* macro_grid = nullptr;
  std::vector<double> qx, qw;
  std::vector<int> cg_iterations;
  std::vector<std::vector<double>> cg_history;  // oracle extension: recursive residual per step
  int num_threads = 1;

  MicroSolver(BioData& d, int n_cells) : data(d), n_cells_axis(n_cells) {}
  void set_grid_locations(const std::vector<std::array<double, 2>>& loc) {  // micro.cpp:398-402
    grid_locations = loc;
    num_grids = static_cast<int>(loc.size());
  }
  void set_macro_solution(const std::vector<double>* u, const std::vector<double>* w, const Grid* mg) {  // :170-175
    sol_u = u;
    sol_w = w;
    macro_grid = mg;
  }
  void setup() {  // micro.cpp:62-69, :72-127
    grid.build(MeshKind::Subdivided, n_cells_axis);
    grid.build_pattern(pattern);
    const int n = grid.n_dofs();
    solutions.assign(num_grids, std::vector<double>(n, 0.0));
    old_solutions.assign(num_grids, std::vector<double>(n, 0.0));
    righthandsides.assign(num_grids, std::vector<double>(n, 0.0));
    system_matrices.assign(num_grids, SparseMatrix());
    cg_iterations.assign(num_grids, 0);
    cg_history.assign(num_grids, {});
    gauss01(fem_q_deg, qx, qw);
  }
  // micro.cpp:183-279 for one (cell, grid)
  void local_assemble(int c, int k, double cell_matrix[4][4], double cell_rhs[4]) const {
    const double k_1 = data.constants.at("kappa_1"), k_2 = data.constants.at("kappa_2");
    const double k_3 = data.constants.at("kappa_3"), k_4 = data.constants.at("kappa_4");
    const double D_2 = data.constants.at("D_2");
    const int p = fem_q_deg;
    const double h = grid.h;
    const double* x = grid_locations[k].data();
    for (int q = 0; q < p * p; ++q) {
      const int qi = q % p, qj = q / p;
      double y[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
      double det_jac, kkt[3];
      data.get_map_data(x, y, det_jac, kkt);
      double g[4][2];
      q1_grads(qx[qi], qx[qj], g);
      const double JxW = qw[qi] * qw[qj] * h * h;
      for (int i = 0; i < 4; ++i) {
        const double gi0 = D_2 * (g[i][0] / h), gi1 = D_2 * (g[i][1] / h);
        const double t0 = gi0 * kkt[0] + gi1 * kkt[1], t1 = gi0 * kkt[1] + gi1 * kkt[2];
        for (int j = 0; j < 4; ++j) cell_matrix[i][j] += (t0 * (g[j][0] / h) + t1 * (g[j][1] / h)) * JxW * det_jac;
      }
    }
    for (int q = 0; q < p * p; ++q) {
      const int qi = q % p, qj = q / p;
      double y[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
      const double det_jac = data.get_map_det_jac(x, y);
      double mapped[2], v[4];
      data.mapping.mmap(x, y, mapped);
      const double rhs_val = data.bulk_rhs_v.mvalue(x, mapped);
      q1_values(qx[qi], qx[qj], v);
      const double JxW = qw[qi] * qw[qj] * h * h;
      for (int i = 0; i < 4; ++i) cell_rhs[i] += rhs_val * v[i] * JxW * det_jac;
    }
    for (int f = 0; f < 4; ++f) {
      if (!grid.face_at_boundary(c, f)) continue;
      for (int q = 0; q < p; ++q) {
        double xi, eta;
        face_point(f, qx[q], xi, eta);
        double y[2] = {grid.cell_x0(c) + h * xi, grid.cell_y0(c) + h * eta};
        const double det_jac = data.get_map_det_jac_bc(x, y, FACE_NORMAL[f]);
        double mp[2], v[4];
        data.mapping.mmap(x, y, mp);
        q1_values(xi, eta, v);
        const double JxW = qw[q] * h;
        for (int i = 0; i < 4; ++i) {
          switch (micro_face_id(f)) {
            case INFLOW_BOUNDARY:
              for (int j = 0; j < 4; ++j) cell_matrix[i][j] += v[i] * det_jac * k_2 * v[j] * JxW;
              cell_rhs[i] += (data.bc_v_1.mvalue(x, mp) + k_1 * (*sol_u)[k]) * v[i] * det_jac * JxW;
              break;
            case OUTFLOW_BOUNDARY:
              for (int j = 0; j < 4; ++j) cell_matrix[i][j] += v[i] * det_jac * k_4 * v[j] * JxW;
              cell_rhs[i] += (data.bc_v_2.mvalue(x, mp) + k_3 * (*sol_w)[k]) * v[i] * det_jac * JxW;
              break;
            case TOP_NEUMANN:
              cell_rhs[i] += data.bc_v_3.mvalue(x, mp) * v[i] * det_jac * JxW;
              break;
            default:
              cell_rhs[i] += data.bc_v_4.mvalue(x, mp) * v[i] * det_jac * JxW;
              break;
          }
        }
      }
    }
  }
  void assemble_one(int k) {
    system_matrices[k].reinit(pattern);
    std::fill(righthandsides[k].begin(), righthandsides[k].end(), 0.0);
    for (int c = 0; c < grid.n_cells(); ++c) {
      double cell_matrix[4][4] = {};
      double cell_rhs[4] = {};
      local_assemble(c, k, cell_matrix, cell_rhs);
      const auto& cd = grid.cell_dofs[c];
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) system_matrices[k].add(cd[i], cd[j], cell_matrix[i][j]);
        righthandsides[k][cd[i]] += cell_rhs[i];
      }
    }
  }
  void solve_one(int k) {  // micro.cpp:309-314
    CGResult r = solve_cg(system_matrices[k], solutions[k], righthandsides[k], 10000, 1e-12, &cg_history[k]);
    if (!r.converged) throw std::runtime_error("bio micro CG: SolverControl::NoConvergence");
    cg_iterations[k] = r.iterations;
  }
  // micro.cpp:292-305 (WorkStream over cells / one task per grid restated as threads over grids;
  // every grid's arithmetic is independent of the schedule)
  void assemble_and_solve_all() {
    parallel_over_grids([this](int k) { assemble_one(k); });
    old_solutions = solutions;
    parallel_over_grids([this](int k) { solve_one(k); });
  }
  template <typename F>
  void parallel_over_grids(F f) {
    if (num_threads <= 1) {
      for (int k = 0; k < num_grids; ++k) f(k);
      return;
    }
    std::vector<std::thread> pool;
    std::vector<std::string> errors(num_threads);
    for (int t = 0; t < num_threads; ++t)
      pool.emplace_back([&, t] {
        try {
          for (int k = t; k < num_grids; k += num_threads) f(k);
        } catch (const std::exception& e) {
          errors[t] = e.what();
        }
      });
    for (auto& th : pool) th.join();
    for (auto& e : errors)
      if (!e.empty()) throw std::runtime_error(e);
  }
  // micro.cpp:381-391: every solution vector becomes the L2 projection of solution_v (QGauss(fem_q_deg))
  void set_exact_solution() {
    for (int k = 0; k < num_grids; ++k) {
      const double* x = grid_locations[k].data();
      solutions[k] = l2_project(grid, pattern, fem_q_deg, [&](const double* y) { return data.solution_v.mvalue(x, y); });
    }
  }
  // micro.cpp:317-338 per grid
  void grid_errors(int k, double& l2, double& h1) const {
    const int p = fem_q_deg;
    const double h = grid.h, fd = 1e-8;
    const double* x = grid_locations[k].data();
    double l2s = 0, h1s = 0;
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      double cl2 = 0, ch1 = 0;
      for (int q = 0; q < p * p; ++q) {
        const int qi = q % p, qj = q / p;
        double v[4], g[4][2];
        q1_values(qx[qi], qx[qj], v);
        q1_grads(qx[qi], qx[qj], g);
        double y[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
        const double JxW = qw[qi] * qw[qj] * h * h;
        double uh = 0, gx = 0, gy = 0;
        for (int i = 0; i < 4; ++i) {
          uh += solutions[k][cd[i]] * v[i];
          gx += solutions[k][cd[i]] * g[i][0] / h;
          gy += solutions[k][cd[i]] * g[i][1] / h;
        }
        const double diff = uh - data.solution_v.mvalue(x, y);
        cl2 += diff * diff * JxW;
        double ex[2];
        for (int dd = 0; dd < 2; ++dd) {
          double q1[2] = {y[0], y[1]}, q2[2] = {y[0], y[1]};
          q1[dd] += fd;
          q2[dd] -= fd;
          ex[dd] = (data.solution_v.mvalue(x, q1) - data.solution_v.mvalue(x, q2)) / (2 * fd);
        }
        ch1 += ((gx - ex[0]) * (gx - ex[0]) + (gy - ex[1]) * (gy - ex[1])) * JxW;
      }
      l2s += cl2;
      h1s += ch1;
    }
    l2 = std::sqrt(l2s);
    h1 = std::sqrt(h1s);
  }
  void compute_all_errors(double& l2_error, double& h1_error) const {  // micro.cpp:317-348
    std::vector<double> el2(num_grids), eh1(num_grids);
    for (int k = 0; k < num_grids; ++k) grid_errors(k, el2[k], eh1[k]);
    l2_error = q1_l2_norm(*macro_grid, el2, fem_q_deg);
    h1_error = q1_l2_norm(*macro_grid, eh1, fem_q_deg);
  }
  double get_residual(int k) const {  // micro.cpp:367-378
    std::vector<double> e(grid.n_dofs());
    for (int i = 0; i < grid.n_dofs(); ++i) e[i] = solutions[k][i] - old_solutions[k][i];
    return q1_l2_norm(grid, e, fem_q_deg);
  }
  void compute_all_residuals(double& l2_residual) const {  // micro.cpp:350-365
    std::vector<double> r(num_grids);
    for (int k = 0; k < num_grids; ++k) r[k] = get_residual(k);
    l2_residual = q1_l2_norm(*macro_grid, r, fem_q_deg);
  }
};

enum { DIRICHLET_BOUNDARY = 0, NEUMANN_BOUNDARY = 1 };

// reference: src/bio-multiscale/macro.{h,cpp}
struct MacroSolver {
  BioData& data;
  int n_cells_axis;
  Grid grid;
  SparsityPattern pattern;
  SparseMatrix system_matrix_u, system_matrix_w;
  std::vector<double> sol_u, sol_w, old_sol_u, old_sol_w, system_rhs_u, system_rhs_w;
  std::vector<double> micro_contribution_u, micro_contribution_w;
  const MicroSolver* micro = nullptr;
  int its_u = 0, its_w = 0;

  MacroSolver(BioData& d, int n_cells) : data(d), n_cells_axis(n_cells) {}
  // macro.cpp:45-58: |x1| = 1 faces are NEUMANN, |x0| = 1 faces DIRICHLET
  static int face_id(int f) { return (f >= 2) ? NEUMANN_BOUNDARY : DIRICHLET_BOUNDARY; }
  void setup() {  // macro.cpp:36-86
    grid.build(MeshKind::Subdivided, n_cells_axis);
    grid.build_pattern(pattern);
    system_matrix_u.reinit(pattern);
    system_matrix_w.reinit(pattern);
    const int n = grid.n_dofs();
    for (auto* v : {&sol_u, &sol_w, &old_sol_u, &old_sol_w, &system_rhs_u, &system_rhs_w, &micro_contribution_u,
                    &micro_contribution_w})
      v->assign(n, 0.0);
  }
  void get_dof_locations(std::vector<std::array<double, 2>>& loc) const {  // macro.cpp:308-315
    loc.resize(grid.n_dofs());
    for (int d = 0; d < grid.n_dofs(); ++d) {
      double p[2];
      grid.dof_point(d, p);
      loc[d] = {p[0], p[1]};
    }
  }
  void set_micro_objects(const MicroSolver* m) { micro = m; }
  // macro.cpp:254-296
  void integrate_micro_cells(int k, const double* macro_point, double& u_contribution, double& w_contribution) const {
    const MicroSolver& ms = *micro;
    u_contribution = 0;
    w_contribution = 0;
    const int p = ms.fem_q_deg;
    const double h = ms.grid.h;
    const double k_2 = data.constants.at("kappa_2"), k_4 = data.constants.at("kappa_4");
    for (int c = 0; c < ms.grid.n_cells(); ++c)
      for (int f = 0; f < 4; ++f) {
        if (!ms.grid.face_at_boundary(c, f)) continue;
        for (int q = 0; q < p; ++q) {
          double xi, eta;
          face_point(f, ms.qx[q], xi, eta);
          const double jxw = ms.qw[q] * h;
          double qp[2] = {ms.grid.cell_x0(c) + h * xi, ms.grid.cell_y0(c) + h * eta};
          double mq[2];
          data.mapping.mmap(macro_point, qp, mq);
          const double det_jac = data.get_map_det_jac_bc(macro_point, qp, FACE_NORMAL[f]);
          const double interp = q1_interp(ms.grid, ms.solutions[k], c, xi, eta);
          switch (micro_face_id(f)) {
            case 0:
              u_contribution += (-k_2 * interp + data.bc_v_1.mvalue(macro_point, mq)) * jxw * det_jac;
              break;
            case 1:
              w_contribution += (-k_4 * interp + data.bc_v_2.mvalue(macro_point, mq)) * jxw * det_jac;
              break;
            default:
              break;
          }
        }
      }
  }
  void compute_microscopic_contribution() {  // macro.cpp:299-306
    for (int i = 0; i < grid.n_dofs(); ++i) {
      double p[2];
      grid.dof_point(i, p);
      integrate_micro_cells(i, p, micro_contribution_u[i], micro_contribution_w[i]);
    }
  }
  // macro.cpp:100-192
  void assemble_system() {
    std::fill(system_matrix_u.val.begin(), system_matrix_u.val.end(), 0.0);
    std::fill(system_matrix_w.val.begin(), system_matrix_w.val.end(), 0.0);
    std::fill(system_rhs_u.begin(), system_rhs_u.end(), 0.0);
    std::fill(system_rhs_w.begin(), system_rhs_w.end(), 0.0);
    compute_microscopic_contribution();
    const int p = 8;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    const double h = grid.h;
    const double k_1 = data.constants.at("kappa_1"), k_3 = data.constants.at("kappa_3"), D_1 = data.constants.at("D_1");
    for (int c = 0; c < grid.n_cells(); ++c) {
      double mu[4][4] = {}, mw[4][4] = {}, ru[4] = {}, rw[4] = {};
      const auto& cd = grid.cell_dofs[c];
      for (int qj = 0; qj < p; ++qj)
        for (int qi = 0; qi < p; ++qi) {
          double v[4], g[4][2];
          q1_values(qx[qi], qx[qj], v);
          q1_grads(qx[qi], qx[qj], g);
          const double JxW = qw[qi] * qw[qj] * h * h;
          double pt[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
          double uc = 0, wc = 0;
          for (int i = 0; i < 4; ++i) {
            uc += micro_contribution_u[cd[i]] * v[i];
            wc += micro_contribution_w[cd[i]] * v[i];
          }
          const double infl = data.inflow_measure.value(pt), outfl = data.outflow_measure.value(pt);
          const double fu = data.bulk_rhs_u.value(pt), fw = data.bulk_rhs_w.value(pt);
          for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 4; ++j) {
              const double laplace_term = ((g[i][0] / h) * (g[j][0] / h) + (g[i][1] / h) * (g[j][1] / h)) * JxW;
              const double mass_term = v[i] * v[j] * JxW;
              mu[i][j] += laplace_term + mass_term * k_1 * infl;
              mw[i][j] += D_1 * laplace_term + mass_term * k_3 * outfl;
            }
            ru[i] += (-uc + fu) * v[i] * JxW;
            rw[i] += (-wc + fw) * v[i] * JxW;
          }
        }
      for (int f = 0; f < 4; ++f) {
        if (!grid.face_at_boundary(c, f)) continue;
        for (int q = 0; q < p; ++q) {
          double xi, eta, v[4];
          face_point(f, qx[q], xi, eta);
          q1_values(xi, eta, v);
          const double JxW = qw[q] * h;
          double pt[2] = {grid.cell_x0(c) + h * xi, grid.cell_y0(c) + h * eta};
          for (int i = 0; i < 4; ++i) {
            if (face_id(f) == NEUMANN_BOUNDARY) {
              ru[i] += v[i] * JxW * data.bc_u_2.value(pt);
              rw[i] += v[i] * JxW * data.bc_w_2.value(pt);
            } else {
              rw[i] += v[i] * JxW * data.bc_w_1.value(pt);
            }
          }
        }
      }
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) {
          system_matrix_u.add(cd[i], cd[j], mu[i][j]);
          system_matrix_w.add(cd[i], cd[j], mw[i][j]);
        }
        system_rhs_u[cd[i]] += ru[i];
        system_rhs_w[cd[i]] += rw[i];
      }
    }
    // Dirichlet for u on the |x0| = 1 faces (all their vertices, corners included)
    std::map<int, double> bv;
    for (int node = 0; node < grid.n_dofs(); ++node) {
      const int ix = node % grid.m;
      if (ix == 0 || ix == grid.m - 1) {
        double pt[2] = {-1.0 + ix * h, -1.0 + (node / grid.m) * h};
        bv[grid.node_dof[node]] = data.bc_u_1.value(pt);
      }
    }
    apply_boundary_values(bv, system_matrix_u, sol_u, system_rhs_u);
  }
  void solve() {  // macro.cpp:194-202
    old_sol_u = sol_u;
    old_sol_w = sol_w;
    CGResult a = solve_cg(system_matrix_u, sol_u, system_rhs_u, 10000, 1e-12);
    CGResult b = solve_cg(system_matrix_w, sol_w, system_rhs_w, 10000, 1e-12);
    if (!a.converged || !b.converged) throw std::runtime_error("bio macro CG: SolverControl::NoConvergence");
    its_u = a.iterations;
    its_w = b.iterations;
  }
  void assemble_and_solve() {  // macro.cpp:317-321
    assemble_system();
    solve();
  }
  static void field_error(const Grid& grid, const std::vector<double>& sol, const MacroFunction& exact, double& l2,
                          double& h1) {
    const int p = 8;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    const double h = grid.h, fd = 1e-8;
    double l2s = 0, h1s = 0;
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      double cl2 = 0, ch1 = 0;
      for (int qj = 0; qj < p; ++qj)
        for (int qi = 0; qi < p; ++qi) {
          double v[4], g[4][2];
          q1_values(qx[qi], qx[qj], v);
          q1_grads(qx[qi], qx[qj], g);
          const double JxW = qw[qi] * qw[qj] * h * h;
          double pt[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
          double uh = 0, gx = 0, gy = 0;
          for (int i = 0; i < 4; ++i) {
            uh += sol[cd[i]] * v[i];
            gx += sol[cd[i]] * g[i][0] / h;
            gy += sol[cd[i]] * g[i][1] / h;
          }
          const double diff = uh - exact.value(pt);
          cl2 += diff * diff * JxW;
          double ex[2];
          for (int dd = 0; dd < 2; ++dd) {
            double q1[2] = {pt[0], pt[1]}, q2[2] = {pt[0], pt[1]};
            q1[dd] += fd;
            q2[dd] -= fd;
            ex[dd] = (exact.value(q1) - exact.value(q2)) / (2 * fd);
          }
          ch1 += ((gx - ex[0]) * (gx - ex[0]) + (gy - ex[1]) * (gy - ex[1])) * JxW;
        }
      l2s += cl2;
      h1s += ch1;
    }
    l2 = std::sqrt(l2s);
    h1 = std::sqrt(h1s);
  }
  void compute_error(double& l2_error, double& h1_error) const {  // macro.cpp:204-228
    double lu, hu, lw, hw;
    field_error(grid, sol_u, data.solution_u, lu, hu);
    field_error(grid, sol_w, data.solution_w, lw, hw);
    l2_error = lu + lw;
    h1_error = hu + hw;
  }
  void compute_residual(double& l2_residual) const {  // macro.cpp:230-245 (the u part is overwritten)
    std::vector<double> e(grid.n_dofs());
    for (int i = 0; i < grid.n_dofs(); ++i) e[i] = sol_w[i] - old_sol_w[i];
    l2_residual = q1_l2_norm(grid, e, 8);
  }
};

struct CycleRecord {
  int cycle;
  double micro_l2 = 0, micro_h1 = 0, macro_l2 = 0, macro_h1 = 0;  // has_solution
  double macro_residual = 0, micro_residual = 0;                  // otherwise
  std::vector<double> sol_u, sol_w, contribution_u, contribution_w;
  std::vector<std::vector<double>> micro_solutions;
  std::vector<int> micro_cg_iterations;
};

// reference: src/bio-multiscale/manager.{h,cpp}
struct Manager {
  BioData data;
  MacroSolver macro_solver;
  MicroSolver micro_solver;
  double error_eps = 1e-2, residual_eps = 1e-5, max_iterations = 1e4;
  double old_residual = 0, residual = 1, old_error = 0, error = 1;
  int cycle = 0;
  int cycle_limit = -1;
  bool keep_solutions = true;
  std::vector<CycleRecord> history;

  // manufactured.cpp:46-48: the manager gets 2^refinement cells per axis
  Manager(int macro_cells, int micro_cells, const std::string& data_file, int num_threads = 1)
      : data(data_file), macro_solver(data, macro_cells), micro_solver(data, micro_cells) {
    micro_solver.num_threads = num_threads;
  }
  void setup() {  // manager.cpp:23-49
    macro_solver.setup();
    std::vector<std::array<double, 2>> loc;
    macro_solver.get_dof_locations(loc);
    micro_solver.set_grid_locations(loc);
    micro_solver.setup();
    micro_solver.set_macro_solution(&macro_solver.sol_u, &macro_solver.sol_w, &macro_solver.grid);
    macro_solver.set_micro_objects(&micro_solver);
    cycle = 0;
    old_residual = 0;
    residual = 1;
    old_error = 0;
    error = 1;
  }
  void fixed_point_iterate() {  // manager.cpp:65-68
    macro_solver.assemble_and_solve();
    micro_solver.assemble_and_solve_all();
  }
  bool check_fixed_point_reached() {  // manager.cpp:70-109
    CycleRecord rec;
    rec.cycle = cycle;
    bool reached;
    if (data.params.get_bool("has_solution")) {
      macro_solver.compute_error(rec.macro_l2, rec.macro_h1);
      micro_solver.compute_all_errors(rec.micro_l2, rec.micro_h1);
      old_error = error;
      error = rec.micro_l2 + rec.macro_l2;
      reached = std::fabs(old_error - error) / error < error_eps;
    } else {
      macro_solver.compute_residual(rec.macro_residual);
      micro_solver.compute_all_residuals(rec.micro_residual);
      old_residual = residual;
      residual = rec.macro_residual + rec.micro_residual;
      reached = std::fabs(old_residual - residual) < residual_eps;
    }
    rec.contribution_u = macro_solver.micro_contribution_u;
    rec.contribution_w = macro_solver.micro_contribution_w;
    rec.micro_cg_iterations = micro_solver.cg_iterations;
    if (keep_solutions) {
      rec.sol_u = macro_solver.sol_u;
      rec.sol_w = macro_solver.sol_w;
      rec.micro_solutions = micro_solver.solutions;
    }
    history.push_back(std::move(rec));
    return reached;
  }
  void run() {  // manager.cpp:51-63
    bool reached = false;
    while (!reached) {
      fixed_point_iterate();
      reached = check_fixed_point_reached();
      cycle++;
      if (cycle > max_iterations) break;
      if (cycle_limit >= 0 && cycle >= cycle_limit) break;
    }
  }
};

}  // namespace bio
}  // namespace orc
