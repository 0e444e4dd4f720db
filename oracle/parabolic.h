// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the reference's elliptic-parabolic driver (solve_parabolic).
// Each function cites the reference lines it follows.
#pragma once
#include <string>
#include <vector>

#include "expr.h"
#include "fem.h"
#include "prm.h"

namespace orc {
namespace parabolic {

// reference: src/tools/pde_data.cpp:110-171 (TwoPressureData<2>)
struct TwoPressureData {
  ParameterHandler params;
  MacroFunction macro_rhs, macro_bc, macro_solution;
  MultiscaleFunction micro_rhs, neumann_bc, robin_bc, micro_solution, init_rho;
  std::map<std::string, double> constants;

  explicit TwoPressureData(const std::string& param_file) {
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
    params.parse_input(param_file);
    for (auto& kv : constants) kv.second = params.get_double(kv.first);
    const std::string mv = "x0,x1,t", msv = "x0,x1,y0,y1,t";
    macro_rhs.initialize(mv, params.get("macro_rhs"), constants, true);
    macro_bc.initialize(mv, params.get("macro_bc"), constants, true);
    macro_solution.initialize(mv, params.get("macro_solution"), constants, true);
    micro_rhs.initialize(msv, params.get("micro_rhs"), constants, nullptr, true);
    neumann_bc.initialize(msv, params.get("micro_bc_neumann"), constants, nullptr, true);
    robin_bc.initialize(msv, params.get("micro_bc_robin"), constants, nullptr, true);
    micro_solution.initialize(msv, params.get("micro_solution"), constants, nullptr, true);
    init_rho.initialize("x0,x1,y0,y1", params.get("init_rho"), constants, nullptr, false);
  }
  void set_time(double time) {  // pde_data.cpp:161-171
    macro_rhs.set_time(time);
    macro_bc.set_time(time);
    macro_solution.set_time(time);
    micro_rhs.set_time(time);
    neumann_bc.set_time(time);
    robin_bc.set_time(time);
    micro_solution.set_time(time);
  }
};

// MatrixCreator::create_mass_matrix / create_laplace_matrix with QGauss<2>(2)
inline void create_mass_and_laplace(const Grid& grid, SparseMatrix& mass, SparseMatrix& laplace) {
  std::vector<double> qx, qw;
  gauss01(2, qx, qw);
  const double h = grid.h;
  for (int c = 0; c < grid.n_cells(); ++c) {
    double m[4][4] = {}, l[4][4] = {};
    for (int qj = 0; qj < 2; ++qj)
      for (int qi = 0; qi < 2; ++qi) {
        double v[4], g[4][2];
        q1_values(qx[qi], qx[qj], v);
        q1_grads(qx[qi], qx[qj], g);
        const double JxW = qw[qi] * qw[qj] * h * h;
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) {
            m[i][j] += v[i] * v[j] * JxW;
            l[i][j] += ((g[i][0] / h) * (g[j][0] / h) + (g[i][1] / h) * (g[j][1] / h)) * JxW;
          }
      }
    const auto& cd = grid.cell_dofs[c];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        mass.add(cd[i], cd[j], m[i][j]);
        laplace.add(cd[i], cd[j], l[i][j]);
      }
  }
}

enum { ROBIN_BOUNDARY = 0, NEUMANN_BOUNDARY = 1 };
// rho_solver.cpp:40-50: faces with |y1| = 1 are NEUMANN, the others ROBIN
inline int micro_face_id(int f) { return f >= 2 ? NEUMANN_BOUNDARY : ROBIN_BOUNDARY; }
inline void face_point(int f, double s, double& xi, double& eta) {
  switch (f) {
    case 0: xi = 0; eta = s; break;
    case 1: xi = 1; eta = s; break;
    case 2: xi = s; eta = 0; break;
    default: xi = s; eta = 1; break;
  }
}

struct PiSolver;

// reference: src/elliptic-parabolic/rho_solver.{h,cpp}
struct RhoSolver {
  TwoPressureData& data;
  int h_inv;
  Grid grid;
  SparsityPattern pattern;
  SparseMatrix mass_matrix, laplace_matrix;
  int num_grids = 1;
  std::vector<std::array<double, 2>> grid_locations;
  std::vector<std::vector<double>> solutions, old_solutions, righthandsides;
  std::vector<SparseMatrix> system_matrices;
  const std::vector<double>*macro_solution = nullptr, *old_macro_solution = nullptr;
  const Grid* macro_grid = nullptr;
  double dt = 0, euler = 1;
  int integration_order = 2;
  std::vector<int> cg_iterations;
  bool share_matrix = true;  // oracle economy: the N system matrices are bit-identical (rho_solver.cpp:148-178)

  RhoSolver(TwoPressureData& d, int h_inv_) : data(d), h_inv(h_inv_) {}
  void set_grid_locations(const std::vector<std::array<double, 2>>& loc) {  // :346-350
    grid_locations = loc;
    num_grids = static_cast<int>(loc.size());
  }
  void set_macro_solutions(const std::vector<double>* s, const std::vector<double>* old, const Grid* mg) {  // :96-102
    macro_solution = s;
    old_macro_solution = old;
    macro_grid = mg;
  }
  // VectorTools::project(dof_handler, constraints, QGauss(3), f, vec): L2 projection
  std::vector<double> project(const MultiscaleFunction& f, const double* x) const {
    const int p = 3;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    const double h = grid.h;
    std::vector<double> rhs(grid.n_dofs(), 0.0), sol(grid.n_dofs(), 0.0);
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      for (int qj = 0; qj < p; ++qj)
        for (int qi = 0; qi < p; ++qi) {
          double v[4];
          q1_values(qx[qi], qx[qj], v);
          double y[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
          const double fv = f.mvalue(x, y);
          for (int i = 0; i < 4; ++i) rhs[cd[i]] += fv * v[i] * qw[qi] * qw[qj] * h * h;
        }
    }
    double nrm = 0;
    for (double r : rhs) nrm += r * r;
    solve_cg(mass_matrix, sol, rhs, 10 * grid.n_dofs() + 100, 1e-12 * std::sqrt(nrm));
    return sol;
  }
  void setup() {  // rho_solver.cpp:28-32, :55-93
    grid.build(MeshKind::Subdivided, h_inv);
    grid.build_pattern(pattern);
    mass_matrix.reinit(pattern);
    laplace_matrix.reinit(pattern);
    create_mass_and_laplace(grid, mass_matrix, laplace_matrix);
    const int n = grid.n_dofs();
    solutions.clear();
    old_solutions.clear();
    for (int k = 0; k < num_grids; ++k) {
      std::vector<double> s = project(data.init_rho, grid_locations[k].data());
      solutions.push_back(s);
      old_solutions.push_back(s);
    }
    righthandsides.assign(num_grids, std::vector<double>(n, 0.0));
    system_matrices.assign(share_matrix ? 1 : num_grids, SparseMatrix());
    cg_iterations.assign(num_grids, 0);
  }
  SparseMatrix& matrix_of(int k) { return system_matrices[share_matrix ? 0 : k]; }
  // rho_solver.cpp:111-218
  void assemble_system() {
    const int p = integration_order;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    const double h = grid.h;
    const double kappa = data.constants.at("kappa"), p_F = data.constants.at("p_F");
    const double R = data.constants.at("R"), D = data.constants.at("D");
    const int n = grid.n_dofs();
    std::vector<double> intermediate(n), rhs_func(n);
    for (int k = 0; k < num_grids; ++k) {
      std::fill(solutions[k].begin(), solutions[k].end(), 0.0);
      mass_matrix.vmult(righthandsides[k], old_solutions[k]);
      laplace_matrix.vmult(intermediate, old_solutions[k]);
      for (int i = 0; i < n; ++i) righthandsides[k][i] += -dt * D * (1 - euler) * intermediate[i];
      // VectorTools::create_right_hand_side(dof, QGauss(2), rhs(x_k, ., t))
      std::fill(rhs_func.begin(), rhs_func.end(), 0.0);
      for (int c = 0; c < grid.n_cells(); ++c) {
        const auto& cd = grid.cell_dofs[c];
        for (int qj = 0; qj < p; ++qj)
          for (int qi = 0; qi < p; ++qi) {
            double v[4];
            q1_values(qx[qi], qx[qj], v);
            double y[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
            const double fv = data.micro_rhs.mvalue(grid_locations[k].data(), y);
            for (int i = 0; i < 4; ++i) rhs_func[cd[i]] += fv * v[i] * qw[qi] * qw[qj] * h * h;
          }
      }
      for (int i = 0; i < n; ++i) righthandsides[k][i] += dt * euler * rhs_func[i];
      if (!share_matrix || k == 0) {
        matrix_of(k).copy_from(mass_matrix);
        matrix_of(k).add_scaled(dt * D * euler, laplace_matrix);
      }
    }
    const double bilin_param = kappa * euler * dt * R;
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      for (int f = 0; f < 4; ++f) {
        if (!grid.face_at_boundary(c, f) || micro_face_id(f) != ROBIN_BOUNDARY) continue;
        double cell_matrix[4][4] = {};
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j)
            for (int q = 0; q < p; ++q) {
              double xi, eta, v[4];
              face_point(f, qx[q], xi, eta);
              q1_values(xi, eta, v);
              cell_matrix[i][j] += bilin_param * v[i] * v[j] * (qw[q] * h);
            }
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j)
            for (int k = 0; k < (share_matrix ? 1 : num_grids); ++k) matrix_of(k).add(cd[i], cd[j], cell_matrix[i][j]);
      }
      for (int k = 0; k < num_grids; ++k)
        for (int f = 0; f < 4; ++f) {
          if (!grid.face_at_boundary(c, f)) continue;
          double cell_rhs[4] = {};
          for (int i = 0; i < 4; ++i)
            for (int q = 0; q < p; ++q) {
              double xi, eta, v[4];
              face_point(f, qx[q], xi, eta);
              q1_values(xi, eta, v);
              double y[2] = {grid.cell_x0(c) + h * xi, grid.cell_y0(c) + h * eta};
              const double dtxvixJxW = dt * v[i] * (qw[q] * h);
              if (micro_face_id(f) == ROBIN_BOUNDARY) {
                const double pi_part = euler * (*macro_solution)[k] + (1 - euler) * (*old_macro_solution)[k];
                const double func_part = euler * data.robin_bc.mvalue(grid_locations[k].data(), y);
                const double old_interp = q1_interp(grid, old_solutions[k], c, xi, eta);
                cell_rhs[i] += (kappa * (pi_part + p_F - R * (1 - euler) * old_interp) + func_part) * dtxvixJxW;
              } else {
                const double func_part = euler * data.neumann_bc.mvalue(grid_locations[k].data(), y);
                cell_rhs[i] += dtxvixJxW * func_part;
              }
            }
          for (int i = 0; i < 4; ++i) righthandsides[k][cd[i]] += cell_rhs[i];
        }
    }
  }
  void solve_time_step() {  // :221-229
    for (int k = 0; k < num_grids; ++k) {
      CGResult r = solve_cg(matrix_of(k), solutions[k], righthandsides[k], 10000, 1e-12);
      if (!r.converged) throw std::runtime_error("rho CG: SolverControl::NoConvergence");
      cg_iterations[k] = r.iterations;
    }
  }
  void iterate(double time_step) {  // :231-237
    dt = time_step;
    assemble_system();
    solve_time_step();
    old_solutions = solutions;
  }
  // rho_solver.cpp:353-359: every solution vector becomes the L2 projection of the exact rho at the current time
  void set_exact_solution() {
    for (int k = 0; k < num_grids; ++k) solutions[k] = project(data.micro_solution, grid_locations[k].data());
  }
  // per-grid part of compute_error (:307-330): L2 and "H1" (integrate_difference with H1_norm,
  // accumulated as an l2 norm over cells) against micro_solution(x_k, ., t), QGauss(3)
  void grid_errors(int k, double& l2, double& h1) const {
    const int p = 3;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    const double h = grid.h, fd = 1e-8;
    const double* x = grid_locations[k].data();
    double l2s = 0, h1s = 0;
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      double cl2 = 0, csemi = 0;
      for (int qj = 0; qj < p; ++qj)
        for (int qi = 0; qi < p; ++qi) {
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
          const double diff = uh - data.micro_solution.mvalue(x, y);
          cl2 += diff * diff * JxW;
          double ex[2];
          for (int dd = 0; dd < 2; ++dd) {
            double q1[2] = {y[0], y[1]}, q2[2] = {y[0], y[1]};
            q1[dd] += fd;
            q2[dd] -= fd;
            ex[dd] = (data.micro_solution.mvalue(x, q1) - data.micro_solution.mvalue(x, q2)) / (2 * fd);
          }
          csemi += ((gx - ex[0]) * (gx - ex[0]) + (gy - ex[1]) * (gy - ex[1])) * JxW;
        }
      l2s += cl2;
      h1s += cl2 + csemi;  // H1_norm per cell: sqrt(L2^2 + seminorm^2)
    }
    l2 = std::sqrt(l2s);
    h1 = std::sqrt(h1s);
  }
  void compute_error(double& l2_error, double& h1_error) const {  // :307-343
    std::vector<double> el2(num_grids), eh1(num_grids);
    for (int k = 0; k < num_grids; ++k) grid_errors(k, el2[k], eh1[k]);
    l2_error = q1_l2_norm(*macro_grid, el2, 3);
    h1_error = q1_l2_norm(*macro_grid, eh1, 3);
  }
};

// reference: src/elliptic-parabolic/pi_solver.{h,cpp}
struct PiSolver {
  TwoPressureData& data;
  int h_inv;
  Grid grid;
  SparsityPattern pattern;
  SparseMatrix system_matrix, laplace_matrix;
  std::vector<double> solution, old_solution, system_rhs, macro_contribution, micro_contribution;
  const RhoSolver* rho = nullptr;
  int count = 0;
  int integration_order = 2;
  int first_step_passes = 0;
  int last_cg_iterations = 0;

  PiSolver(TwoPressureData& d, int h_inv_) : data(d), h_inv(h_inv_) {}
  void setup() {  // pi_solver.cpp:28-59
    grid.build(MeshKind::Subdivided, h_inv);
    grid.build_pattern(pattern);
    system_matrix.reinit(pattern);
    laplace_matrix.reinit(pattern);
    SparseMatrix dummy;
    dummy.reinit(pattern);
    create_mass_and_laplace(grid, dummy, laplace_matrix);
    const int n = grid.n_dofs();
    solution.assign(n, 0.0);
    old_solution.assign(n, 5.0);
    system_rhs.assign(n, 0.0);
    macro_contribution.assign(n, 0.0);
    micro_contribution.assign(n, 0.0);
  }
  void set_micro_solutions(const RhoSolver* r) { rho = r; }
  void get_dof_locations(std::vector<std::array<double, 2>>& loc) const {
    loc.resize(grid.n_dofs());
    for (int d = 0; d < grid.n_dofs(); ++d) {
      double p[2];
      grid.dof_point(d, p);
      loc[d] = {p[0], p[1]};
    }
  }
  double get_micro_mass(int k) const {  // pi_solver.cpp:177-197
    const RhoSolver& r = *rho;
    std::vector<double> qx, qw;
    gauss01(integration_order, qx, qw);
    const double h = r.grid.h;
    double integral = 0;
    for (int c = 0; c < r.grid.n_cells(); ++c)
      for (int qj = 0; qj < integration_order; ++qj)
        for (int qi = 0; qi < integration_order; ++qi)
          integral += q1_interp(r.grid, r.solutions[k], c, qx[qi], qx[qj]) * (qw[qi] * qw[qj] * h * h);
    return integral;
  }
  void get_microscopic_contribution(std::vector<double>& out, bool nonlinear) const {  // :235-246
    for (int i = 0; i < grid.n_dofs(); ++i)
      out[i] = nonlinear ? std::fmin(std::fabs(get_micro_mass(i)), 1.0) : get_micro_mass(i);
  }
  void get_pi_contribution_rhs(const std::vector<double>& pi, std::vector<double>& out, bool nonlinear) const {  // :61-73
    const double theta = data.constants.at("theta");
    for (size_t i = 0; i < pi.size(); ++i) {
      if (nonlinear) {
        const double a = std::fabs(pi[i]);
        out[i] = theta * std::fmin(a, std::sqrt(a));
      } else {
        out[i] = theta * pi[i];
      }
    }
  }
  void assemble_system() {  // :75-122
    std::vector<double> qx, qw;
    gauss01(integration_order, qx, qw);
    const double h = grid.h;
    std::fill(system_matrix.val.begin(), system_matrix.val.end(), 0.0);
    std::fill(system_rhs.begin(), system_rhs.end(), 0.0);
    system_matrix.add_scaled(data.constants.at("A"), laplace_matrix);
    const bool nl = data.params.get_bool("nonlinear");
    get_microscopic_contribution(micro_contribution, nl);
    get_pi_contribution_rhs(old_solution, macro_contribution, nl);
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      double cell_rhs[4] = {};
      for (int qj = 0; qj < integration_order; ++qj)
        for (int qi = 0; qi < integration_order; ++qi) {
          double v[4];
          q1_values(qx[qi], qx[qj], v);
          double pt[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
          double pi_q = 0, rho_q = 0;
          for (int i = 0; i < 4; ++i) {
            pi_q += macro_contribution[cd[i]] * v[i];
            rho_q += micro_contribution[cd[i]] * v[i];
          }
          const double functional = rho_q * pi_q + data.macro_rhs.value(pt);
          for (int i = 0; i < 4; ++i) cell_rhs[i] += v[i] * functional * (qw[qi] * qw[qj] * h * h);
        }
      for (int i = 0; i < 4; ++i) system_rhs[cd[i]] += cell_rhs[i];
    }
    std::map<int, double> bv;
    for (int node = 0; node < grid.n_dofs(); ++node)
      if (grid.node_on_boundary(node)) {
        double pt[2] = {-1.0 + (node % grid.m) * h, -1.0 + (node / grid.m) * h};
        bv[grid.node_dof[node]] = data.macro_bc.value(pt);
      }
    apply_boundary_values(bv, system_matrix, solution, system_rhs);
  }
  void solve() {  // :143-149
    CGResult r = solve_cg(system_matrix, solution, system_rhs, 10000, 1e-12);
    if (!r.converged) throw std::runtime_error("pi CG: SolverControl::NoConvergence");
    last_cg_iterations = r.iterations;
  }
  double solution_difference() const {  // :267-277
    std::vector<double> d(solution.size());
    for (size_t i = 0; i < d.size(); ++i) d[i] = solution[i] - old_solution[i];
    return q1_l2_norm(grid, d, 3);
  }
  void iterate() {  // :248-265
    assemble_system();
    solve();
    if (count == 0) {
      double diff = 1;
      while (diff > 1e-9) {
        diff = solution_difference();
        old_solution = solution;
        assemble_system();
        solve();
        ++first_step_passes;
      }
    }
    count++;
    old_solution = solution;
  }
  void compute_error(double& l2, double& h1) const {  // :151-168
    const int p = 3;
    std::vector<double> qx, qw;
    gauss01(p, qx, qw);
    const double h = grid.h, fd = 1e-8;
    double l2s = 0, h1s = 0;
    for (int c = 0; c < grid.n_cells(); ++c) {
      const auto& cd = grid.cell_dofs[c];
      for (int qj = 0; qj < p; ++qj)
        for (int qi = 0; qi < p; ++qi) {
          double v[4], g[4][2];
          q1_values(qx[qi], qx[qj], v);
          q1_grads(qx[qi], qx[qj], g);
          double pt[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
          const double JxW = qw[qi] * qw[qj] * h * h;
          double uh = 0, gx = 0, gy = 0;
          for (int i = 0; i < 4; ++i) {
            uh += solution[cd[i]] * v[i];
            gx += solution[cd[i]] * g[i][0] / h;
            gy += solution[cd[i]] * g[i][1] / h;
          }
          const double diff = uh - data.macro_solution.value(pt);
          l2s += diff * diff * JxW;
          double ex[2];
          for (int dd = 0; dd < 2; ++dd) {
            double q1[2] = {pt[0], pt[1]}, q2[2] = {pt[0], pt[1]};
            q1[dd] += fd;
            q2[dd] -= fd;
            ex[dd] = (data.macro_solution.value(q1) - data.macro_solution.value(q2)) / (2 * fd);
          }
          h1s += ((gx - ex[0]) * (gx - ex[0]) + (gy - ex[1]) * (gy - ex[1])) * JxW;
        }
    }
    l2 = std::sqrt(l2s);
    h1 = std::sqrt(h1s);
  }
};

struct StepRecord {
  int it;
  double time, micro_l2, macro_l2, micro_h1, macro_h1;
  std::vector<double> pi;
  std::vector<double> micro_mass;
  std::vector<std::vector<double>> rho;
  std::vector<int> micro_cg_iterations;
};

// reference: src/elliptic-parabolic/time_manager.{h,cpp}
struct TimeManager {
  TwoPressureData data;
  PiSolver pi_solver;
  RhoSolver rho_solver;
  double time_step, final_time = 0.5, time = 0;
  int it = 0;
  int step_limit = -1;
  bool keep_solutions = true;
  std::vector<StepRecord> history;

  TimeManager(int macro_h_inv, int micro_h_inv, int t_inv, const std::string& data_file)
      : data(data_file), pi_solver(data, macro_h_inv), rho_solver(data, micro_h_inv), time_step(1.0 / t_inv) {}
  void setup() {  // time_manager.cpp:19-39
    pi_solver.setup();
    std::vector<std::array<double, 2>> loc;
    pi_solver.get_dof_locations(loc);
    rho_solver.set_grid_locations(loc);
    rho_solver.setup();
    rho_solver.set_macro_solutions(&pi_solver.solution, &pi_solver.solution, &pi_solver.grid);
    pi_solver.set_micro_solutions(&rho_solver);
    it = 0;
  }
  void iterate() {  // :65-70
    data.set_time(time);
    pi_solver.iterate();
    rho_solver.iterate(time_step);
  }
  void run() {  // :41-62
    while (time < final_time) {
      time += time_step;
      it++;
      iterate();
      StepRecord r;
      r.it = it;
      r.time = time;
      pi_solver.compute_error(r.macro_l2, r.macro_h1);
      rho_solver.compute_error(r.micro_l2, r.micro_h1);
      r.micro_cg_iterations = rho_solver.cg_iterations;
      r.micro_mass = pi_solver.micro_contribution;
      if (keep_solutions) {
        r.pi = pi_solver.solution;
        r.rho = rho_solver.solutions;
      }
      history.push_back(std::move(r));
      if (step_limit >= 0 && it >= step_limit) break;
    }
  }
};

}  // namespace parabolic
}  // namespace orc
