// solve_elliptic: elliptic-elliptic two-scale driver on top of the GPU micro path.
// Host mirror of reference src/elliptic-elliptic/{macro,manager}.{h,cpp} and manufactured.cpp.
#pragma once
#include <cstdio>
#include <fstream>

#include "micro_solver.h"
#include "output.h"

namespace ms {
namespace elliptic {

struct Functions {
  EllipticData data;
  MacroFunction macro_rhs, macro_bc, macro_solution;
  explicit Functions(const std::string& file) : data(file) {
    macro_rhs.initialize(data.params.get("macro_rhs"), data.constants);
    macro_bc.initialize(data.params.get("macro_bc"), data.constants);
    macro_solution.initialize(data.params.get("macro_solution"), data.constants);
  }
};

// reference: macro.{h,cpp}
class MacroSolver {
 public:
  MacroSolver(Functions& f, unsigned int refine_level) : pde_data(f), refine_level(refine_level) {}
  void setup() {  // macro.cpp:19-43
    mesh.build(MSGPU_MESH_REFINE_GLOBAL, 1 << refine_level);
    system_matrix.reinit(mesh);
    solution.assign(mesh.n, 0.0);
    system_rhs.assign(mesh.n, 0.0);
    micro_contribution.assign(mesh.n, 0.0);
  }
  void get_dof_locations(std::vector<std::array<double, 2>>& loc) const {  // macro.cpp:168-175
    loc.resize(mesh.n);
    for (int d = 0; d < mesh.n; ++d) loc[d] = {mesh.x_of(d), mesh.y_of(d)};
  }
  void set_micro_objects(MicroSolver* m) { micro = m; }  // macro.cpp:128-131
  // get_micro_bulk for every macro DoF (macro.cpp:134-166) is one coupling kernel on the device
  void compute_microscopic_contribution() { micro->coupling(micro_contribution); }
  // the macro matrix never changes and the load is affine in the micro contribution,
  // so after one host assembly (contribution = 0) the system lives on the device.
  void enable_device() {
    std::fill(micro_contribution.begin(), micro_contribution.end(), 0.0);
    assemble_system(false);
    CsrMatrix C = mass_operator(mesh, 8, 1.0);
    zero_rows(C, boundary_values);
    micro->macro_setup(1, MSGPU_MESH_REFINE_GLOBAL, 1 << refine_level);
    micro->macro_set_system(0, 0, mesh, system_matrix, &C, system_rhs, solution);
    on_device = true;
  }
  void assemble_system(bool fetch_contribution = true) {  // macro.cpp:57-106
    Quadrature Q(8);
    if (fetch_contribution) compute_microscopic_contribution();
    system_matrix.zero();
    std::fill(system_rhs.begin(), system_rhs.end(), 0.0);
    const double h = mesh.h;
    for (size_t c = 0; c < mesh.cell_dofs.size(); ++c) {
      const auto& cd = mesh.cell_dofs[c];
      const double ox = -1.0 + mesh.cell_cx[c] * h, oy = -1.0 + mesh.cell_cy[c] * h;
      double A[4][4] = {}, b[4] = {};
      for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 8; ++i) {
          double v[4], g[4][2];
          shape(Q.pt[i], Q.pt[j], v);
          shape_grad(Q.pt[i], Q.pt[j], h, g);
          const double jxw = Q.wt[i] * Q.wt[j] * h * h;
          double contribution = 0;
          for (int a = 0; a < 4; ++a) contribution += micro_contribution[cd[a]] * v[a];
          const double f = pde_data.macro_rhs.value(ox + h * Q.pt[i], oy + h * Q.pt[j]);
          for (int a = 0; a < 4; ++a) {
            for (int bb = 0; bb < 4; ++bb) A[a][bb] += (g[a][0] * g[bb][0] + g[a][1] * g[bb][1]) * jxw;
            b[a] += v[a] * (contribution + f) * jxw;
          }
        }
      for (int a = 0; a < 4; ++a) {
        for (int bb = 0; bb < 4; ++bb) system_matrix.add(cd[a], cd[bb], A[a][bb]);
        system_rhs[cd[a]] += b[a];
      }
    }
    boundary_values.clear();
    for (int d = 0; d < mesh.n; ++d)
      if (mesh.on_boundary(d)) boundary_values[d] = pde_data.macro_bc.value(mesh.x_of(d), mesh.y_of(d));
    apply_boundary_values(boundary_values, system_matrix, solution, system_rhs);
  }
  void solve() {  // macro.cpp:108-114
    if (micro->solves_macro_here()) last_iterations = solve_cg(system_matrix, solution, system_rhs, 1000, 1e-12);
    micro->share_macro_solution(solution);
    std::printf("\t %d CG iterations to convergence (macro)\n", last_iterations);
  }
  void run() {  // macro.cpp:177-181
    if (on_device) {
      compute_microscopic_contribution();  // leaves the integrals on the device; the host copy is for the records
      micro->assemble_begin();  // matrices, loads and weights of the next micro run: beside the macro solve
      if (micro->solves_macro_here()) {
        last_iterations = micro->macro_solve(1e-12, 1000, true)[0];
        micro->macro_get_solution(0, solution);
      }
      micro->share_macro_solution(solution);
      std::printf("\t %d CG iterations to convergence (macro)\n", last_iterations);
      return;
    }
    assemble_system();
    solve();
  }
  void compute_error(double& l2, double& h1) const { field_errors(mesh, solution, pde_data.macro_solution, 8, l2, h1); }

  MacroMesh mesh;
  std::vector<double> solution, micro_contribution;
  int last_iterations = 0;

 private:
  Functions& pde_data;
  unsigned int refine_level;
  CsrMatrix system_matrix;
  std::vector<double> system_rhs;
  std::map<int, double> boundary_values;
  MicroSolver* micro = nullptr;
  bool on_device = false;
};

struct CycleRecord {
  int cycle;
  double mL2, mH1, ML2, MH1;
  std::vector<double> macro_solution, micro_contribution, micro_solutions;
  std::vector<int> micro_iterations;
};

// reference: manager.{h,cpp}
class Manager {
 public:
  Manager(unsigned int macro_refinement, unsigned int micro_refinement, const std::string& data_file,
          const std::string& output_file, int device = -1)
      : functions(data_file),
        macro_solver(functions, macro_refinement),
        micro_solver(MSGPU_ELLIPTIC, 1 << micro_refinement, 12, MSGPU_MESH_REFINE_GLOBAL, device),
        ct_file_name(output_file) {
    std::printf("Running elliptic-elliptic solver with data from %s, storing results in %s\n", data_file.c_str(),
                output_file.c_str());
    const auto& p = functions.data.params;
    const Constants& c = functions.data.constants;
    micro_solver.set_function(MSGPU_FN_MAPPING, p.get("mapping"), c);
    micro_solver.set_function(MSGPU_FN_JACOBIAN, p.get("jac_mapping"), c);
    micro_solver.set_function(MSGPU_FN_RHS, p.get("micro_rhs"), c);
    micro_solver.set_function(MSGPU_FN_BC, p.get("micro_bc"), c);
    micro_solver.set_function(MSGPU_FN_SOLUTION, p.get("micro_solution"), c);
  }
  void setup() {  // manager.cpp:17-35
    macro_solver.setup();
    std::vector<std::array<double, 2>> dof_locations;
    macro_solver.get_dof_locations(dof_locations);
    micro_solver.set_grid_locations(dof_locations);
    micro_solver.setup();
    micro_solver.set_macro_solution(&macro_solver.solution);
    macro_solver.set_micro_objects(&micro_solver);
    if (device_macro_enabled()) macro_solver.enable_device();
    if (device_macro_enabled()) micro_solver.norm_setup(macro_solver.mesh);  // the stop test on the device as well
    cycle = 0;
  }
  void run() {  // manager.cpp:37-51
    double old_residual = 1, residual = 0;
    while (std::fabs(old_residual - residual) > eps) {
      fixed_point_iterate();
      compute_errors(old_residual, residual);
      std::printf("Old residual %.2e, new residual %.2e\n", old_residual, residual);
      cycle++;
      if (cycle > max_iterations) {
        std::printf("Can't get the residual small enough...\n");
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
  void output_results() {  // manager.cpp:79-88; several ranks: rank 0 writes the table and the macro file
    if (ct_file_name.empty()) return;
    if (distributed().rank == 0) {
      std::ofstream out(ct_file_name, std::ios::app);
      out << convergence_table();
    }
    if (solution_files_enabled()) patch_and_write_solutions();
  }
  // manager.cpp:91-131.  NB micro-error.gpl holds (computed - projected exact) of the middle grid and, like the
  // reference, this leaves the PROJECTED EXACT solutions in the micro solver.  The rank that owns the middle
  // grid writes the three micro files.
  void patch_and_write_solutions() {
    ensure_directory(output_dir);
    if (distributed().rank == 0)
      write_gnuplot_file(output_dir + "/macro-solution.gpl", macro_solver.mesh, macro_solver.solution, "solution");
    const int some_int = static_cast<int>(micro_solver.get_num_grids() / 2);
    if (!micro_solver.owns(some_int)) return;
    MacroMesh micro_mesh;
    micro_mesh.build(MSGPU_MESH_REFINE_GLOBAL, micro_solver.cells_per_axis());
    const std::vector<double> computed = micro_solver.solution(some_int);
    write_gnuplot_file(output_dir + "/micro-computed.gpl", micro_mesh, computed, "solution");
    std::printf("Exact micro-solution set\n");
    micro_solver.set_exact_solution();
    const std::vector<double> exact = micro_solver.solution(some_int);
    std::vector<double> error(computed);
    for (size_t i = 0; i < error.size(); ++i) error[i] -= exact[i];
    write_gnuplot_file(output_dir + "/micro-error.gpl", micro_mesh, error, "solution");
    write_gnuplot_file(output_dir + "/micro-exact.gpl", micro_mesh, exact, "solution");
  }

  Functions functions;
  MacroSolver macro_solver;
  MicroSolver micro_solver;
  double eps = 1e-4, max_iterations = 1e4;
  int cycle_limit = -1;      // extension for K-cycle parity runs on the diverging parameter sets
  bool keep_solutions = false;
  std::string output_dir = "results";  // the reference writes into results/ of the working directory
  std::vector<CycleRecord> history;

 private:
  int cycle = 0;
  std::string ct_file_name;
  void fixed_point_iterate() {  // manager.cpp:53-58
    macro_solver.run();
    micro_solver.run();
    if (!micro_solver.iterations.empty())  // the last grid of this rank (micro.cpp:181 prints the last grid's count)
      std::printf("\t %d CG iterations to convergence (micro)\n", micro_solver.iterations.back());
  }
  void compute_errors(double& old_residual, double& residual) {  // manager.cpp:60-77
    CycleRecord r;
    r.cycle = cycle;
    macro_solver.compute_error(r.ML2, r.MH1);
    micro_solver.compute_all_errors(macro_solver.mesh, 12, r.mL2, r.mH1);
    r.micro_contribution = macro_solver.micro_contribution;
    r.micro_iterations = micro_solver.iterations;
    if (keep_solutions) {
      r.macro_solution = macro_solver.solution;
      micro_solver.all_solutions(r.micro_solutions);
    }
    old_residual = residual;
    std::printf("Macro residual: %.2e\tMicro residual: %.2e\n", r.ML2, r.mL2);
    residual = r.mL2 + r.ML2;
    history.push_back(std::move(r));
  }
};

}  // namespace elliptic
}  // namespace ms
