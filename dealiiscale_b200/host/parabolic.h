// solve_parabolic: elliptic-parabolic two-scale driver on top of the GPU micro path.
// Host mirror of reference src/elliptic-parabolic/{pi_solver,time_manager}.{h,cpp} and two_scale.cpp.
#pragma once
#include <cmath>
#include <cstdio>
#include <fstream>

#include "micro_solver.h"
#include "output.h"

namespace ms {
namespace parabolic {

struct Functions {
  TwoPressureData data;
  MacroFunction macro_rhs, macro_bc, macro_solution;
  explicit Functions(const std::string& file) : data(file) {
    macro_rhs.initialize(data.params.get("macro_rhs"), data.constants, true);
    macro_bc.initialize(data.params.get("macro_bc"), data.constants, true);
    macro_solution.initialize(data.params.get("macro_solution"), data.constants, true);
  }
  void set_time(double t) {  // pde_data.cpp:161-171 (the micro functions get t with every time step call)
    macro_rhs.set_time(t);
    macro_bc.set_time(t);
    macro_solution.set_time(t);
  }
};

// reference: pi_solver.{h,cpp}
class PiSolver {
 public:
  PiSolver(Functions& f, unsigned int h_inv) : pde_data(f), h_inv(h_inv) {}
  void setup() {  // pi_solver.cpp:28-59
    mesh.build(MSGPU_MESH_SUBDIVIDED, h_inv);
    system_matrix.reinit(mesh);
    laplace_matrix.reinit(mesh);
    Quadrature Q(2);
    for (size_t c = 0; c < mesh.cell_dofs.size(); ++c) {
      double L[4][4] = {};
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
          double g[4][2];
          shape_grad(Q.pt[i], Q.pt[j], mesh.h, g);
          for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) L[a][b] += (g[a][0] * g[b][0] + g[a][1] * g[b][1]) * Q.wt[i] * Q.wt[j] * mesh.h * mesh.h;
        }
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) laplace_matrix.add(mesh.cell_dofs[c][a], mesh.cell_dofs[c][b], L[a][b]);
    }
    solution.assign(mesh.n, 0.0);
    old_solution.assign(mesh.n, 5.0);
    system_rhs.assign(mesh.n, 0.0);
    macro_contribution.assign(mesh.n, 0.0);
    micro_contribution.assign(mesh.n, 0.0);
  }
  void get_dof_locations(std::vector<std::array<double, 2>>& loc) const {
    loc.resize(mesh.n);
    for (int d = 0; d < mesh.n; ++d) loc[d] = {mesh.x_of(d), mesh.y_of(d)};
  }
  void set_micro_solutions(MicroSolver* m) { micro = m; }
  // get_micro_mass for every macro DoF (pi_solver.cpp:177-197, :235-246): one coupling kernel
  void get_microscopic_contribution(std::vector<double>& out, bool nonlinear) {
    micro->coupling(out);
    if (nonlinear)
      for (double& v : out) v = std::fmin(std::fabs(v), 1.0);
  }
  void get_pi_contribution_rhs(const std::vector<double>& pi, std::vector<double>& out, bool nonlinear) const {
    const double theta = pde_data.data.constants.at("theta");
    for (size_t i = 0; i < pi.size(); ++i) {
      const double a = std::fabs(pi[i]);
      out[i] = nonlinear ? theta * std::fmin(a, std::sqrt(a)) : theta * pi[i];
    }
  }
  void assemble_system() {  // pi_solver.cpp:75-122
    Quadrature Q(2);
    const double h = mesh.h;
    system_matrix.zero();
    std::fill(system_rhs.begin(), system_rhs.end(), 0.0);
    const double A = pde_data.data.constants.at("A");
    for (size_t i = 0; i < system_matrix.val.size(); ++i) system_matrix.val[i] += A * laplace_matrix.val[i];
    const bool nl = pde_data.data.params.get_bool("nonlinear");
    get_microscopic_contribution(micro_contribution, nl);
    get_pi_contribution_rhs(old_solution, macro_contribution, nl);
    for (size_t c = 0; c < mesh.cell_dofs.size(); ++c) {
      const auto& cd = mesh.cell_dofs[c];
      const double ox = -1.0 + mesh.cell_cx[c] * h, oy = -1.0 + mesh.cell_cy[c] * h;
      double b[4] = {};
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
          double v[4];
          shape(Q.pt[i], Q.pt[j], v);
          double pi_q = 0, rho_q = 0;
          for (int a = 0; a < 4; ++a) {
            pi_q += macro_contribution[cd[a]] * v[a];
            rho_q += micro_contribution[cd[a]] * v[a];
          }
          // device path: the product of the two finite-element functions is added by k_macro_product (macro.cu)
          const double functional = (device_product ? 0.0 : rho_q * pi_q) +
                                    pde_data.macro_rhs.value(ox + h * Q.pt[i], oy + h * Q.pt[j]);
          for (int a = 0; a < 4; ++a) b[a] += v[a] * functional * Q.wt[i] * Q.wt[j] * h * h;
        }
      for (int a = 0; a < 4; ++a) system_rhs[cd[a]] += b[a];
    }
    std::map<int, double> boundary_values;
    for (int d = 0; d < mesh.n; ++d)
      if (mesh.on_boundary(d)) boundary_values[d] = pde_data.macro_bc.value(mesh.x_of(d), mesh.y_of(d));
    apply_boundary_values(boundary_values, system_matrix, solution, system_rhs);
  }
  // A * Laplace with Dirichlet elimination is constant.  Of the load the host assembles what
  // depends on t only (macro_rhs and the boundary values, :97, :102-105); the product of the two finite-element
  // functions rho_h pi_h (:92-97) is evaluated on the device from the device-resident micro masses and the previous
  // solution.  Every rank solves (MSGPU_MACRO_BROADCAST: rank 0 alone - then the product stays on the host as well).
  void enable_device() {
    assemble_system();
    const std::vector<double> zero(mesh.n, 0.0);
    micro->macro_setup(1, MSGPU_MESH_SUBDIVIDED, h_inv);
    micro->macro_set_system(0, 0, mesh, system_matrix, nullptr, zero, solution);
    on_device = true;
    const char* host_product = std::getenv("MSGPU_HOST_PRODUCT");  // A/B: keep the product term on the host
    if (!MicroSolver::macro_broadcast() && !(host_product && host_product[0] == '1')) {
      std::vector<unsigned char> dirichlet(mesh.n, 0);
      for (int d = 0; d < mesh.n; ++d) dirichlet[d] = mesh.on_boundary(d) ? 1 : 0;
      micro->macro_set_product_load(0, pde_data.data.constants.at("theta"), pde_data.data.params.get_bool("nonlinear"),
                                    old_solution, dirichlet);
      device_product = true;
    }
  }
  void solve() {
    if (micro->solves_macro_here()) {
      if (on_device) {
        micro->macro_set_solution(0, solution);  // boundary values may move with t (apply_boundary_values set them)
        micro->macro_set_rhs(0, system_rhs);
        last_iterations = micro->macro_solve(1e-12, 10000, false)[0];
        micro->macro_get_solution(0, solution);
      } else {
        last_iterations = solve_cg(system_matrix, solution, system_rhs, 10000, 1e-12);
      }
    }
    micro->share_macro_solution(solution);
  }
  double solution_difference() const {  // pi_solver.cpp:267-277
    std::vector<double> d(mesh.n);
    for (int i = 0; i < mesh.n; ++i) d[i] = solution[i] - old_solution[i];
    return nodal_l2_norm(mesh, d, 3);
  }
  void iterate() {  // pi_solver.cpp:248-265
    assemble_system();
    solve();
    if (count == 0) {
      double diff = 1;
      while (diff > 1e-9) {
        std::printf("Difference: %.3e, reiterating first step\n", diff);
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
  void compute_error(double& l2, double& h1) const { field_errors(mesh, solution, pde_data.macro_solution, 3, l2, h1); }

  MacroMesh mesh;
  std::vector<double> solution, old_solution, micro_contribution;
  int first_step_passes = 0, last_iterations = 0;

 private:
  Functions& pde_data;
  unsigned int h_inv;
  CsrMatrix system_matrix, laplace_matrix;
  std::vector<double> system_rhs, macro_contribution;
  MicroSolver* micro = nullptr;
  int count = 0;
  bool on_device = false, device_product = false;
};

struct StepRecord {
  int it;
  double time, mL2, ML2, mH1, MH1;
  std::vector<double> pi, micro_mass, rho;
  std::vector<int> micro_iterations;
};

// reference: time_manager.{h,cpp}
class TimeManager {
 public:
  TimeManager(unsigned int macro_h_inv, unsigned int micro_h_inv, unsigned int t_inv, const std::string& data_file,
              const std::string& out_file, int device = -1)
      : functions(data_file),
        pi_solver(functions, macro_h_inv),
        rho_solver(MSGPU_PARABOLIC, micro_h_inv, 2, MSGPU_MESH_SUBDIVIDED, device),
        time_step(1.0 / t_inv),
        ct_file_name(out_file) {
    std::printf("Using a time step of %.2e\n", time_step);
    const auto& p = functions.data.params;
    const Constants& c = functions.data.constants;
    rho_solver.set_function(MSGPU_FN_RHS, p.get("micro_rhs"), c, true);
    rho_solver.set_function(MSGPU_FN_SOLUTION, p.get("micro_solution"), c, true);
    rho_solver.set_function(MSGPU_FN_BC_V1, p.get("micro_bc_robin"), c, true);
    rho_solver.set_function(MSGPU_FN_BC_V2, p.get("micro_bc_neumann"), c, true);
    rho_solver.set_function(MSGPU_FN_INIT, p.get("init_rho"), c, false);
    rho_solver.set_scalars(functions.data.scalars());
  }
  void setup() {  // time_manager.cpp:19-39
    pi_solver.setup();
    std::vector<std::array<double, 2>> dof_locations;
    pi_solver.get_dof_locations(dof_locations);
    rho_solver.set_grid_locations(dof_locations);
    rho_solver.setup();
    rho_solver.set_macro_solution(&pi_solver.solution);
    pi_solver.set_micro_solutions(&rho_solver);
    if (device_macro_enabled()) pi_solver.enable_device();
    if (device_macro_enabled()) rho_solver.norm_setup(pi_solver.mesh);  // the stop test on the device as well
    it = 0;
  }
  void run() {  // time_manager.cpp:41-62
    while (time < final_time) {
      time += time_step;
      it++;
      std::printf("\nSolving for t = %f\n", time);
      functions.set_time(time);  // iterate(): time_manager.cpp:65-70
      pi_solver.iterate();
      rho_solver.iterate(time_step, time);
      write_plot();
      StepRecord r;
      r.it = it;
      r.time = time;
      pi_solver.compute_error(r.ML2, r.MH1);
      rho_solver.compute_all_errors(pi_solver.mesh, 3, r.mL2, r.mH1);
      r.micro_iterations = rho_solver.iterations;
      r.micro_mass = pi_solver.micro_contribution;
      if (keep_solutions) {
        r.pi = pi_solver.solution;
        rho_solver.all_solutions(r.rho);
      }
      history.push_back(std::move(r));
      if (step_limit >= 0 && it >= step_limit) break;
    }
    output_results();
  }
  std::string convergence_table() const {
    std::vector<std::vector<std::string>> rows;
    char tbuf[32];
    for (const auto& r : history) {
      std::snprintf(tbuf, sizeof tbuf, "%g", r.time);
      rows.push_back({std::to_string(r.it), tbuf, std::to_string(pi_solver.mesh.cell_dofs.size()),
                      std::to_string(pi_solver.mesh.n), sci3(r.mL2), sci3(r.ML2), sci3(r.mH1), sci3(r.MH1)});
    }
    return format_table({"iteration", "time", "cells", "dofs", "mL2", "ML2", "mH1", "MH1"}, rows);
  }
  // time_manager.cpp:92-109, called after every time step (:69); grid 0 and the macro solution live on rank 0
  void write_plot() {
    if (ct_file_name.empty() || !solution_files_enabled() || distributed().rank != 0) return;
    ensure_directory(output_dir);
    write_vtk_file(output_dir + "/micro-solution" + int_to_string(it, 3) + ".vtk", micro_mesh(), rho_solver.solution(0),
                   "solution");
    write_vtk_file(output_dir + "/macro-solution" + int_to_string(it, 3) + ".vtk", pi_solver.mesh, pi_solver.solution,
                   "solution");
  }
  void output_results() {  // time_manager.cpp:111-143
    if (ct_file_name.empty()) return;
    const bool root = distributed().rank == 0;
    if (root) {
      std::ofstream out(ct_file_name, std::ios::app);
      out << convergence_table();
    }
    if (!solution_files_enabled()) return;
    ensure_directory(output_dir);
    if (root)
      write_gnuplot_file(output_dir + "/final_macro_solution.gpl", pi_solver.mesh, pi_solver.solution, "solution");
    const unsigned int index = rho_solver.get_num_grids() / 2;
    if (rho_solver.owns(index))
      write_gnuplot_file(output_dir + "/final_micro_solution.gpl", micro_mesh(), rho_solver.solution(index), "solution");
    std::vector<std::array<double, 2>> dof_locations;
    pi_solver.get_dof_locations(dof_locations);
    patch_micro_solutions(dof_locations);
  }
  // RhoSolver::patch_micro_solutions (rho_solver.cpp:240-283): every micro solution drawn on a small copy of the
  // micro mesh centred at its macro point, all into one gnuplot file (several ranks: one file per rank,
  // patched_micro_solution.<rank>.gpl, each holding the grids that rank owns)
  void patch_micro_solutions(const std::vector<std::array<double, 2>>& locations) {
    const Distributed& d = distributed();
    std::ofstream output(output_dir + (d.active() ? "/patched_micro_solution." + std::to_string(d.rank) + ".gpl"
                                                  : std::string("/patched_micro_solution.gpl")));
    const double side_length = std::pow(static_cast<double>(locations.size()), 0.5);
    const double size = 1.0 / (static_cast<int>(side_length) - 1);  // get_micro_grid_size
    const double micro_size = 2 * size;
    const MacroMesh& mm = micro_mesh();
    std::vector<double> all;
    rho_solver.all_solutions(all);  // this rank's grids
    const size_t n = rho_solver.n_dofs(), first = rho_solver.first_grid();
    std::vector<double> one(n);
    for (size_t i = 0; i < rho_solver.get_num_local_grids(); ++i) {
      std::copy(all.begin() + i * n, all.begin() + (i + 1) * n, one.begin());
      Placement pl;
      pl.x0 = -0.5 * micro_size + locations[first + i][0];
      pl.y0 = -0.5 * micro_size + locations[first + i][1];
      pl.width = micro_size;
      write_gnuplot(output, mm, one, "solution", pl);
    }
  }
  const MacroMesh& micro_mesh() {
    if (micro_mesh_.n == 0) micro_mesh_.build(MSGPU_MESH_SUBDIVIDED, rho_solver.cells_per_axis());
    return micro_mesh_;
  }

  Functions functions;
  PiSolver pi_solver;
  MicroSolver rho_solver;
  double time_step, final_time = 0.5, time = 0;
  int step_limit = -1;
  bool keep_solutions = false;
  std::string output_dir = "results";
  std::vector<StepRecord> history;

 private:
  int it = 0;
  std::string ct_file_name;
  MacroMesh micro_mesh_;
};

}  // namespace parabolic
}  // namespace ms
