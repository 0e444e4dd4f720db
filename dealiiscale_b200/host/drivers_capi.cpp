// C entry points around the C++ host drivers (libmsdrivers.so) so that the parity tests can run
// whole fixed-point loops / time loops and read back their per-cycle history.
#include <cstring>
#include <string>

#include "bio.h"
#include "elliptic.h"
#include "parabolic.h"

namespace {
thread_local std::string g_error;
template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}
void copy(const std::vector<double>& v, double* out) {
  if (out) std::copy(v.begin(), v.end(), out);
}
void copy(const std::vector<int>& v, int* out) {
  if (out) std::copy(v.begin(), v.end(), out);
}
}  // namespace

extern "C" {

const char* msdrv_last_error() { return g_error.c_str(); }

// ---- the out-of-band exchange of the multi-process drivers (distributed.h): host-only self test --------
// Reads RANK / WORLD_SIZE / MSGPU_RENDEZVOUS_DIR; every rank contributes a blob derived from its rank and
// checks what it gets back from allgather, broadcast and barrier.  Returns the world size, -1 on a mismatch.
int msdrv_rendezvous_selftest(int rounds) {
  return guarded([&] {
    ms::Distributed& d = ms::distributed();
    for (int it = 0; it < rounds; ++it) {
      unsigned char mine[128];
      for (int i = 0; i < 128; ++i) mine[i] = static_cast<unsigned char>(7 * d.rank + 3 * i + it);
      std::vector<unsigned char> all;
      d.allgather(mine, sizeof mine, all);
      for (int r = 0; r < d.world; ++r)
        for (int i = 0; i < 128; ++i)
          if (all[128 * r + i] != static_cast<unsigned char>(7 * r + 3 * i + it)) return -1;
      unsigned char id[64];
      for (int i = 0; i < 64; ++i) id[i] = d.rank == 1 % d.world ? static_cast<unsigned char>(200 - i + it) : 0;
      d.broadcast(id, sizeof id, 1 % d.world);
      for (int i = 0; i < 64; ++i)
        if (id[i] != static_cast<unsigned char>(200 - i + it)) return -1;
      d.barrier();
    }
    d.finalize();
    return d.world;
  });
}

// Distributed::usable_ranks: the largest rank count <= world for which no block of msgpu_partition is empty
int msdrv_usable_ranks(int n_total, int world) { return ms::Distributed::usable_ranks(n_total, world); }

// ---- solution files (output.h): host-only, no device involved -----------------------------------
// format 0 = DataOut::write_gnuplot, 1 = DataOut::write_vtk; nodal[(nc+1)^2] in the reference numbering of the mesh kind
int msdrv_write_solution(const char* path, int format, int mesh_kind, int cells_per_axis, const double* nodal,
                         const char* name) {
  return guarded([&] {
    ms::MacroMesh mesh;
    mesh.build(mesh_kind, cells_per_axis);
    const std::vector<double> v(nodal, nodal + mesh.n);
    if (format == 0) ms::write_gnuplot_file(path, mesh, v, name);
    else ms::write_vtk_file(path, mesh, v, name);
    return 0;
  });
}
int msdrv_write_grid_locations(const char* path, const double* xy, int n) {
  return guarded([&] {
    std::vector<std::array<double, 2>> loc(n);
    for (int i = 0; i < n; ++i) loc[i] = {xy[2 * i], xy[2 * i + 1]};
    ms::write_grid_locations(path, loc, n);
    return 0;
  });
}
// support points of the mesh in its reference numbering, xy[(nc+1)^2][2]
int msdrv_mesh_points(int mesh_kind, int cells_per_axis, double* xy) {
  return guarded([&] {
    ms::MacroMesh mesh;
    mesh.build(mesh_kind, cells_per_axis);
    for (int d = 0; d < mesh.n; ++d) {
      xy[2 * d] = mesh.x_of(d);
      xy[2 * d + 1] = mesh.y_of(d);
    }
    return 0;
  });
}

// ---- solve_elliptic -----------------------------------------------------------------------------
void* msdrv_elliptic_create(const char* prm, int macro_ref, int micro_ref) {
  try {
    auto* m = new ms::elliptic::Manager(macro_ref, micro_ref, prm, "");
    m->keep_solutions = true;
    m->setup();
    return m;
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}
void msdrv_elliptic_destroy(void* h) { delete static_cast<ms::elliptic::Manager*>(h); }
int msdrv_elliptic_run(void* h, int cycle_limit) {
  return guarded([&] {
    auto* m = static_cast<ms::elliptic::Manager*>(h);
    m->cycle_limit = cycle_limit;
    m->run();
    return static_cast<int>(m->history.size());
  });
}
void msdrv_elliptic_sizes(void* h, int* N, int* n) {
  auto* m = static_cast<ms::elliptic::Manager*>(h);
  *N = m->macro_solver.mesh.n;
  *n = m->micro_solver.n_dofs();
}
void msdrv_elliptic_history(void* h, int cycle, double* errs4, double* macro, double* contribution, double* micro,
                            int* its) {
  const auto& r = static_cast<ms::elliptic::Manager*>(h)->history.at(cycle);
  errs4[0] = r.mL2;
  errs4[1] = r.mH1;
  errs4[2] = r.ML2;
  errs4[3] = r.MH1;
  copy(r.macro_solution, macro);
  copy(r.micro_contribution, contribution);
  copy(r.micro_solutions, micro);
  copy(r.micro_iterations, its);
}
int msdrv_elliptic_table(void* h, char* buf, int len) {
  const std::string t = static_cast<ms::elliptic::Manager*>(h)->convergence_table();
  std::strncpy(buf, t.c_str(), len - 1);
  buf[len - 1] = 0;
  return static_cast<int>(t.size());
}

// ---- solve_biomath ------------------------------------------------------------------------------
void* msdrv_bio_create(const char* prm, int macro_cells, int micro_cells) {
  try {
    auto* m = new ms::bio::Manager(macro_cells, micro_cells, prm, "");
    m->keep_solutions = true;
    m->setup();
    return m;
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}
void msdrv_bio_destroy(void* h) { delete static_cast<ms::bio::Manager*>(h); }
int msdrv_bio_run(void* h, int cycle_limit) {
  return guarded([&] {
    auto* m = static_cast<ms::bio::Manager*>(h);
    m->cycle_limit = cycle_limit;
    m->run();
    return static_cast<int>(m->history.size());
  });
}
void msdrv_bio_sizes(void* h, int* N, int* n) {
  auto* m = static_cast<ms::bio::Manager*>(h);
  *N = m->macro_solver.mesh.n;
  *n = m->micro_solver.n_dofs();
}
void msdrv_bio_history(void* h, int cycle, double* scalars6, double* sol_u, double* sol_w, double* cu, double* cw,
                       double* micro, int* its) {
  const auto& r = static_cast<ms::bio::Manager*>(h)->history.at(cycle);
  scalars6[0] = r.mL2;
  scalars6[1] = r.mH1;
  scalars6[2] = r.ML2;
  scalars6[3] = r.MH1;
  scalars6[4] = r.macro_residual;
  scalars6[5] = r.micro_residual;
  copy(r.sol_u, sol_u);
  copy(r.sol_w, sol_w);
  copy(r.contribution_u, cu);
  copy(r.contribution_w, cw);
  copy(r.micro_solutions, micro);
  copy(r.micro_iterations, its);
}

// ---- solve_parabolic ----------------------------------------------------------------------------
void* msdrv_parabolic_create(const char* prm, int macro_h_inv, int micro_h_inv, int t_inv) {
  try {
    auto* m = new ms::parabolic::TimeManager(macro_h_inv, micro_h_inv, t_inv, prm, "");
    m->keep_solutions = true;
    m->setup();
    return m;
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}
void msdrv_parabolic_destroy(void* h) { delete static_cast<ms::parabolic::TimeManager*>(h); }
int msdrv_parabolic_run(void* h, int step_limit) {
  return guarded([&] {
    auto* m = static_cast<ms::parabolic::TimeManager*>(h);
    m->step_limit = step_limit;
    m->run();
    return static_cast<int>(m->history.size());
  });
}
void msdrv_parabolic_sizes(void* h, int* N, int* n) {
  auto* m = static_cast<ms::parabolic::TimeManager*>(h);
  *N = m->pi_solver.mesh.n;
  *n = m->rho_solver.n_dofs();
}
int msdrv_parabolic_first_step_passes(void* h) {
  return static_cast<ms::parabolic::TimeManager*>(h)->pi_solver.first_step_passes;
}
void msdrv_parabolic_history(void* h, int step, double* scalars5, double* pi, double* mass, double* rho, int* its) {
  const auto& r = static_cast<ms::parabolic::TimeManager*>(h)->history.at(step);
  scalars5[0] = r.time;
  scalars5[1] = r.mL2;
  scalars5[2] = r.ML2;
  scalars5[3] = r.mH1;
  scalars5[4] = r.MH1;
  copy(r.pi, pi);
  copy(r.micro_mass, mass);
  copy(r.rho, rho);
  copy(r.micro_iterations, its);
}

}  // extern "C"
