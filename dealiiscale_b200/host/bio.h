// solve_biomath: bio-multiscale driver on top of the GPU micro path.
// Host mirror of reference src/bio-multiscale/{macro,manager}.{h,cpp} and manufactured.cpp.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <fstream>

#include "micro_solver.h"
#include "output.h"

namespace ms {
namespace bio {

struct Functions {
  BioData data;
  MacroFunction solution_u, bulk_rhs_u, bc_u_1, bc_u_2, inflow_measure, outflow_measure;
  MacroFunction solution_w, bulk_rhs_w, bc_w_1, bc_w_2;
  explicit Functions(const std::string& file) : data(file) {
    auto init = [&](MacroFunction& f, const char* key) { f.initialize(data.params.get(key), data.constants); };
    init(solution_u, "solution_u"); init(bulk_rhs_u, "bulk_rhs_u"); init(bc_u_1, "bc_u_1"); init(bc_u_2, "bc_u_2");
    init(inflow_measure, "inflow_measure"); init(outflow_measure, "outflow_measure");
    init(solution_w, "solution_w"); init(bulk_rhs_w, "bulk_rhs_w"); init(bc_w_1, "bc_w_1"); init(bc_w_2, "bc_w_2");
  }
};

// reference: macro.{h,cpp}
class MacroSolver {
 public:
  MacroSolver(Functions& f, unsigned int n_cells) : pde_data(f), n_cells(n_cells) {}
  void setup() {  // macro.cpp:36-86: |x1| = 1 faces NEUMANN, |x0| = 1 faces DIRICHLET
    mesh.build(MSGPU_MESH_SUBDIVIDED, n_cells);
    system_matrix_u.reinit(mesh);
    system_matrix_w.reinit(mesh);
    for (auto* v : {&sol_u, &sol_w, &old_sol_u, &old_sol_w, &system_rhs_u, &system_rhs_w, &micro_contribution_u,
                    &micro_contribution_w})
      v->assign(mesh.n, 0.0);
  }
  void get_dof_locations(std::vector<std::array<double, 2>>& loc) const {
    loc.resize(mesh.n);
    for (int d = 0; d < mesh.n; ++d) loc[d] = {mesh.x_of(d), mesh.y_of(d)};
  }
  void set_micro_objects(MicroSolver* m) { micro = m; }
  // integrate_micro_cells for every macro DoF (macro.cpp:254-306): one coupling kernel
  void compute_microscopic_contribution() { micro->coupling(micro_contribution_u, &micro_contribution_w); }
  // both macro matrices are constant and both loads affine in the flux integrals
  // ((-uc + fu), (-wc + fw) below), so after one host assembly the two systems live on the device.
  void enable_device() {
    std::fill(micro_contribution_u.begin(), micro_contribution_u.end(), 0.0);
    std::fill(micro_contribution_w.begin(), micro_contribution_w.end(), 0.0);
    assemble_system(false);
    CsrMatrix Cu = mass_operator(mesh, 8, -1.0), Cw = mass_operator(mesh, 8, -1.0);
    zero_rows(Cu, boundary_values);
    micro->macro_setup(2, MSGPU_MESH_SUBDIVIDED, n_cells);
    micro->macro_set_system(0, 0, mesh, system_matrix_u, &Cu, system_rhs_u, sol_u);
    micro->macro_set_system(1, 1, mesh, system_matrix_w, &Cw, system_rhs_w, sol_w);
    on_device = true;
  }
  void assemble_system(bool fetch_contribution = true) {  // macro.cpp:100-192
    system_matrix_u.zero();
    system_matrix_w.zero();
    std::fill(system_rhs_u.begin(), system_rhs_u.end(), 0.0);
    std::fill(system_rhs_w.begin(), system_rhs_w.end(), 0.0);
    if (fetch_contribution) compute_microscopic_contribution();
    Quadrature Q(8);
    const double h = mesh.h;
    const double k_1 = pde_data.data.constants.at("kappa_1"), k_3 = pde_data.data.constants.at("kappa_3");
    const double D_1 = pde_data.data.constants.at("D_1");
    for (size_t c = 0; c < mesh.cell_dofs.size(); ++c) {
      const auto& cd = mesh.cell_dofs[c];
      const int cx = mesh.cell_cx[c], cy = mesh.cell_cy[c];
      const double ox = -1.0 + cx * h, oy = -1.0 + cy * h;
      double Au[4][4] = {}, Aw[4][4] = {}, bu[4] = {}, bw[4] = {};
      for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 8; ++i) {
          double v[4], g[4][2];
          shape(Q.pt[i], Q.pt[j], v);
          shape_grad(Q.pt[i], Q.pt[j], h, g);
          const double jxw = Q.wt[i] * Q.wt[j] * h * h;
          const double x0 = ox + h * Q.pt[i], x1 = oy + h * Q.pt[j];
          double uc = 0, wc = 0;
          for (int a = 0; a < 4; ++a) {
            uc += micro_contribution_u[cd[a]] * v[a];
            wc += micro_contribution_w[cd[a]] * v[a];
          }
          const double infl = pde_data.inflow_measure.value(x0, x1), outfl = pde_data.outflow_measure.value(x0, x1);
          const double fu = pde_data.bulk_rhs_u.value(x0, x1), fw = pde_data.bulk_rhs_w.value(x0, x1);
          for (int a = 0; a < 4; ++a) {
            for (int b = 0; b < 4; ++b) {
              const double laplace_term = (g[a][0] * g[b][0] + g[a][1] * g[b][1]) * jxw;
              const double mass_term = v[a] * v[b] * jxw;
              Au[a][b] += laplace_term + mass_term * k_1 * infl;
              Aw[a][b] += D_1 * laplace_term + mass_term * k_3 * outfl;
            }
            bu[a] += (-uc + fu) * v[a] * jxw;
            bw[a] += (-wc + fw) * v[a] * jxw;
          }
        }
      // boundary faces: 0 x-min, 1 x-max (DIRICHLET id), 2 y-min, 3 y-max (NEUMANN id)
      const bool at[4] = {cx == 0, cx == (int)n_cells - 1, cy == 0, cy == (int)n_cells - 1};
      for (int f = 0; f < 4; ++f) {
        if (!at[f]) continue;
        for (int q = 0; q < 8; ++q) {
          const double s = Q.pt[q];
          const double xi = f == 0 ? 0 : f == 1 ? 1 : s, eta = f == 2 ? 0 : f == 3 ? 1 : s;
          double v[4];
          shape(xi, eta, v);
          const double jxw = Q.wt[q] * h, x0 = ox + h * xi, x1 = oy + h * eta;
          for (int a = 0; a < 4; ++a) {
            if (f >= 2) {
              bu[a] += v[a] * jxw * pde_data.bc_u_2.value(x0, x1);
              bw[a] += v[a] * jxw * pde_data.bc_w_2.value(x0, x1);
            } else {
              bw[a] += v[a] * jxw * pde_data.bc_w_1.value(x0, x1);
            }
          }
        }
      }
      for (int a = 0; a < 4; ++a) {
        for (int b = 0; b < 4; ++b) {
          system_matrix_u.add(cd[a], cd[b], Au[a][b]);
          system_matrix_w.add(cd[a], cd[b], Aw[a][b]);
        }
        system_rhs_u[cd[a]] += bu[a];
        system_rhs_w[cd[a]] += bw[a];
      }
    }
    boundary_values.clear();  // u only, on |x0| = 1 (macro.cpp:189-191)
    for (int d = 0; d < mesh.n; ++d)
      if (mesh.dof_ix[d] == 0 || mesh.dof_ix[d] == mesh.m - 1)
        boundary_values[d] = pde_data.bc_u_1.value(mesh.x_of(d), mesh.y_of(d));
    apply_boundary_values(boundary_values, system_matrix_u, sol_u, system_rhs_u);
  }
  void solve() {  // macro.cpp:194-202
    old_sol_u = sol_u;
    old_sol_w = sol_w;
    if (micro->solves_macro_here()) {
      solve_cg(system_matrix_u, sol_u, system_rhs_u, 10000, 1e-12);
      solve_cg(system_matrix_w, sol_w, system_rhs_w, 10000, 1e-12);
    }
    micro->share_macro_solution(sol_u, &sol_w);
  }
  void assemble_and_solve() {
    if (on_device) {
      const char* tm = std::getenv("MSGPU_DRIVER_TIMING");
      const bool fine = tm && tm[0] == '2';
      const auto t0 = std::chrono::steady_clock::now();
      compute_microscopic_contribution();  // integrals stay on the device; host copies are for the records
      // the matrices, loads and weights of the micro systems do not depend on what is solved for next: their assembly
      // runs on a second stream beside the macro solve (the reference does one after the other, manager.cpp:65-68)
      micro->assemble_begin();
      old_sol_u = sol_u;
      old_sol_w = sol_w;
      const auto t1 = std::chrono::steady_clock::now();
      std::vector<int> its;
      if (micro->solves_macro_here()) {
        its = micro->macro_solve(1e-12, 10000, true);
        const auto t2 = std::chrono::steady_clock::now();
        micro->macro_get_solution(0, sol_u);
        micro->macro_get_solution(1, sol_w);
        if (fine) {
          const auto t3 = std::chrono::steady_clock::now();
          auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
          std::printf("[timing2] macro: integrals to host %.3f ms, rhs + CG (%d, %d steps) + adopt %.3f ms, solutions to host %.3f ms\n",
                      ms(t0, t1), its.size() > 0 ? its[0] : 0, its.size() > 1 ? its[1] : 0, ms(t1, t2), ms(t2, t3));
        }
      }
      micro->share_macro_solution(sol_u, &sol_w);
      return;
    }
    assemble_system();
    solve();
  }
  void compute_error(double& l2_error, double& h1_error) const {  // macro.cpp:204-228
    double lu, hu, lw, hw;
    field_errors(mesh, sol_u, pde_data.solution_u, 8, lu, hu);
    field_errors(mesh, sol_w, pde_data.solution_w, 8, lw, hw);
    l2_error = lu + lw;
    h1_error = hu + hw;
  }
  void compute_residual(double& l2_residual) const {  // macro.cpp:230-245: the u part is overwritten by the w part
    std::vector<double> e(mesh.n);
    for (int i = 0; i < mesh.n; ++i) e[i] = sol_w[i] - old_sol_w[i];
    l2_residual = (micro && micro->norms_on_device()) ? micro->norm_of_values(e, 8) : nodal_l2_norm(mesh, e, 8);
  }

  MacroMesh mesh;
  std::vector<double> sol_u, sol_w, old_sol_u, old_sol_w, micro_contribution_u, micro_contribution_w;

 private:
  Functions& pde_data;
  unsigned int n_cells;
  CsrMatrix system_matrix_u, system_matrix_w;
  std::vector<double> system_rhs_u, system_rhs_w;
  std::map<int, double> boundary_values;
  MicroSolver* micro = nullptr;
  bool on_device = false;
};

struct CycleRecord {
  int cycle;
  double mL2 = 0, mH1 = 0, ML2 = 0, MH1 = 0, macro_residual = 0, micro_residual = 0;
  std::vector<double> sol_u, sol_w, contribution_u, contribution_w, micro_solutions;
  std::vector<int> micro_iterations;
};

// reference: manager.{h,cpp}
class Manager {
 public:
  Manager(unsigned int macro_cells, unsigned int micro_cells, const std::string& data_file,
          const std::string& out_file, int num_threads = 0, int device = -1)
      : functions(data_file),
        macro_solver(functions, macro_cells),
        micro_solver(MSGPU_BIO, micro_cells, 12, MSGPU_MESH_SUBDIVIDED, device),
        num_threads(num_threads),
        ct_file_name(out_file) {
    std::printf("Running elliptic-elliptic solver with data from %s, storing results in %s\n", data_file.c_str(),
                out_file.c_str());
    const auto& p = functions.data.params;
    const Constants& c = functions.data.constants;
    micro_solver.set_function(MSGPU_FN_MAPPING, p.get("mapping"), c);
    micro_solver.set_function(MSGPU_FN_JACOBIAN, p.get("jac_mapping"), c);
    micro_solver.set_function(MSGPU_FN_RHS, p.get("bulk_rhs_v"), c);
    micro_solver.set_function(MSGPU_FN_SOLUTION, p.get("solution_v"), c);
    micro_solver.set_function(MSGPU_FN_BC_V1, p.get("bc_v_1"), c);
    micro_solver.set_function(MSGPU_FN_BC_V2, p.get("bc_v_2"), c);
    micro_solver.set_function(MSGPU_FN_BC_V3, p.get("bc_v_3"), c);
    micro_solver.set_function(MSGPU_FN_BC_V4, p.get("bc_v_4"), c);
    micro_solver.set_scalars(functions.data.scalars());
  }
  void setup() {  // manager.cpp:23-49
    macro_solver.setup();
    std::vector<std::array<double, 2>> dof_locations;
    macro_solver.get_dof_locations(dof_locations);
    micro_solver.set_grid_locations(dof_locations);
    micro_solver.setup();
    micro_solver.set_macro_solution(&macro_solver.sol_u, &macro_solver.sol_w);
    macro_solver.set_micro_objects(&micro_solver);
    if (device_macro_enabled()) macro_solver.enable_device();
    if (device_macro_enabled()) micro_solver.norm_setup(macro_solver.mesh);  // the stop test on the device as well
    cycle = 0;
    old_residual = 0;
    residual = 1;
    old_error = 0;
    error = 1;
  }
  void run() {  // manager.cpp:51-63
    bool reached = false;
    while (!reached) {
      fixed_point_iterate();
      check_fixed_point_reached(reached);
      cycle++;
      if (cycle > max_iterations) {
        std::printf("Residual too high after %d iterations, stopping\n", cycle);
        break;
      }
      if (cycle_limit >= 0 && cycle >= cycle_limit) break;
    }
    output_results();
  }
  std::string convergence_table() const {
    std::vector<std::vector<std::string>> rows;
    for (const auto& r : history)
      rows.push_back({std::to_string(r.cycle), std::to_string(macro_solver.mesh.cell_dofs.size()),
                      std::to_string(macro_solver.mesh.n), sci3(r.mL2), sci3(r.mH1), sci3(r.ML2), sci3(r.MH1)});
    return format_table({"cycle", "cells", "dofs", "mL2", "mH1", "ML2", "MH1"}, rows);
  }
  void output_results() {  // manager.cpp:111-123; several ranks: rank 0 writes the table and the macro files
    if (ct_file_name.empty()) return;
    if (functions.data.params.get_bool("has_solution") && distributed().rank == 0) {
      std::ofstream out(ct_file_name, std::ios::app);
      out << convergence_table();
    }
    if (solution_files_enabled()) patch_and_write_solutions();
  }
  // manager.cpp:125-178: one VTK file per micro grid, as the reference writes them; every rank writes the
  // grids it owns (numbered globally), the owner of the middle grid writes micro-error / micro-exact
  void patch_and_write_solutions() {
    ensure_directory(output_dir + "/micro-solutions");
    if (distributed().rank == 0) {
      write_vtk_file(output_dir + "/u-solution.vtk", macro_solver.mesh, macro_solver.sol_u, "u");
      write_vtk_file(output_dir + "/w-solution.vtk", macro_solver.mesh, macro_solver.sol_w, "w");
      write_micro_grid_locations(output_dir + "/micro-solutions/grid_locations.txt");
    }
    MacroMesh micro_mesh;
    micro_mesh.build(MSGPU_MESH_SUBDIVIDED, micro_solver.cells_per_axis());
    const unsigned int first = micro_solver.first_grid(), last = first + micro_solver.get_num_local_grids();
    const int n = micro_solver.n_dofs();
    const int some_int = static_cast<int>(micro_solver.get_num_grids() / 2);
    std::vector<double> computed_mid;
    const unsigned int batch = 256;  // grids fetched from the device per copy
    std::vector<double> buf;
    const VtkSeriesWriter series(micro_mesh, "v");  // 4225 files of 4096 cells at `6 6`: formatted by several threads
    for (unsigned int k0 = first; k0 < last; k0 += batch) {
      const unsigned int k1 = std::min(last, k0 + batch);
      micro_solver.solutions_range(k0, k1, buf);
      std::vector<std::string> paths;
      for (unsigned int k = k0; k < k1; ++k) {
        if (static_cast<int>(k) == some_int)
          computed_mid.assign(buf.begin() + static_cast<size_t>(k - k0) * n, buf.begin() + static_cast<size_t>(k - k0 + 1) * n);
        paths.push_back(output_dir + "/micro-solutions/v-solution-" + std::to_string(k) + ".vtk");
      }
      series.write(paths, buf.data(), n);
    }
    if (functions.data.params.get_bool("has_solution") && micro_solver.owns(some_int)) {
      std::printf("Exact micro-solution set\n");
      micro_solver.set_exact_solution();
      const std::vector<double> exact = micro_solver.solution(some_int);
      std::vector<double> error(computed_mid);
      for (size_t i = 0; i < error.size(); ++i) error[i] -= exact[i];
      write_vtk_file(output_dir + "/micro-error.vtk", micro_mesh, error, "v");
      write_vtk_file(output_dir + "/micro-exact.vtk", micro_mesh, exact, "v");
    }
  }
  void write_micro_grid_locations(const std::string& filename) {  // manager.cpp:181-195
    std::printf("Writing grid locations to file\n");
    std::vector<std::array<double, 2>> grid_locations;
    macro_solver.get_dof_locations(grid_locations);
    write_grid_locations(filename, grid_locations, micro_solver.get_num_grids());
  }

  Functions functions;
  MacroSolver macro_solver;
  MicroSolver micro_solver;
  int num_threads;
  double old_residual = 0, residual = 1, old_error = 0, error = 1;
  double error_eps = 1e-2, residual_eps = 1e-5, max_iterations = 1e4;
  int cycle_limit = -1;
  bool keep_solutions = false;
  std::string output_dir = "results";
  std::vector<CycleRecord> history;

 private:
  int cycle = 0;
  std::string ct_file_name;
  void fixed_point_iterate() {  // manager.cpp:65-68
    const bool timing = std::getenv("MSGPU_DRIVER_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    macro_solver.assemble_and_solve();
    const auto t1 = std::chrono::steady_clock::now();
    micro_solver.assemble_and_solve_all();
    const auto t2 = std::chrono::steady_clock::now();
    if (timing) {
      long long its = 0;
      for (int v : micro_solver.iterations) its += v;
      std::printf("[timing] macro %.2f ms, micro %.2f ms (mean CG steps %.1f)\n",
                  std::chrono::duration<double, std::milli>(t1 - t0).count(),
                  std::chrono::duration<double, std::milli>(t2 - t1).count(),
                  micro_solver.iterations.empty() ? 0.0 : double(its) / micro_solver.iterations.size());
    }
  }
  void check_fixed_point_reached(bool& reached) {  // manager.cpp:70-109
    const auto t_check = std::chrono::steady_clock::now();
    struct Report {
      std::chrono::steady_clock::time_point t0;
      ~Report() {
        if (std::getenv("MSGPU_DRIVER_TIMING"))
          std::printf("[timing] stop test %.2f ms\n",
                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
    } report{t_check};
    CycleRecord r;
    r.cycle = cycle;
    if (functions.data.params.get_bool("has_solution")) {
      macro_solver.compute_error(r.ML2, r.MH1);
      micro_solver.compute_all_errors(macro_solver.mesh, 12, r.mL2, r.mH1);
      old_error = error;
      std::printf("Macro error: %.2e\tMicro error: %.2e\n", r.ML2, r.mL2);
      error = r.mL2 + r.ML2;
      std::printf("Old error %.2e, new error %.2e\n", old_error, error);
      reached = std::fabs(old_error - error) / error < error_eps;
    } else {
      const auto ts0 = std::chrono::steady_clock::now();
      macro_solver.compute_residual(r.macro_residual);
      const auto ts1 = std::chrono::steady_clock::now();
      micro_solver.compute_all_residuals(macro_solver.mesh, 12, r.micro_residual);
      if (const char* tm = std::getenv("MSGPU_DRIVER_TIMING"); tm && tm[0] == '2')
        std::printf("[timing2] stop test: macro residual %.3f ms, micro residuals %.3f ms\n",
                    std::chrono::duration<double, std::milli>(ts1 - ts0).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts1).count());
      old_residual = residual;
      std::printf("Macro residual: %.2e\tMicro residual: %.2e\n", r.macro_residual, r.micro_residual);
      residual = r.macro_residual + r.micro_residual;
      std::printf("Old residual %.2e, new residual %.2e\n", old_residual, residual);
      reached = std::fabs(old_residual - residual) < residual_eps;
    }
    r.contribution_u = macro_solver.micro_contribution_u;
    r.contribution_w = macro_solver.micro_contribution_w;
    r.micro_iterations = micro_solver.iterations;
    if (keep_solutions) {
      r.sol_u = macro_solver.sol_u;
      r.sol_w = macro_solver.sol_w;
      micro_solver.all_solutions(r.micro_solutions);
    }
    history.push_back(std::move(r));
  }
};

}  // namespace bio
}  // namespace ms
