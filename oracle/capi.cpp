// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// C entry points so that tests/ and bench.py's cpu_baseline / --impl reference legs can
// drive the CPU restatement through ctypes.  Nothing under dealiiscale_b200/ may load this.
#include <chrono>
#include <cstring>
#include <string>
#include <thread>

#include "bio.h"
#include "elliptic.h"
#include "mapmap.h"
#include "parabolic.h"

using namespace orc;

namespace {
thread_local std::string g_err;
template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
void dense_from(const SparseMatrix& A, double* out) {
  const int n = A.sp->n;
  std::fill(out, out + static_cast<size_t>(n) * n, 0.0);
  for (int r = 0; r < n; ++r)
    for (int k = A.sp->rowstart[r]; k < A.sp->rowstart[r + 1]; ++k) out[static_cast<size_t>(r) * n + A.sp->col[k]] = A.val[k];
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- expression evaluator ------------------------------------------------------------------
// vars: comma separated names; returns 0 and the value, or -1.
int orc_eval(const char* expr, const char* vars, const char* const* cnames, const double* cvals, int nc,
             const double* values, double* out) {
  return guarded([&] {
    std::vector<std::string> v;
    std::string cur;
    for (const char* p = vars; *p; ++p) {
      if (*p == ',') {
        v.push_back(cur);
        cur.clear();
      } else if (*p != ' ') {
        cur.push_back(*p);
      }
    }
    if (!cur.empty()) v.push_back(cur);
    std::map<std::string, double> c;
    for (int i = 0; i < nc; ++i) c[cnames[i]] = cvals[i];
    Expression e(expr, v, c);
    *out = e.eval(values);
    return 0;
  });
}

void orc_gauss01(int p, double* x, double* w) {
  std::vector<double> xv, wv;
  gauss01(p, xv, wv);
  std::memcpy(x, xv.data(), sizeof(double) * p);
  std::memcpy(w, wv.data(), sizeof(double) * p);
}

// dof numbering: node_dof[iy*m+ix] for mesh kind (0 refine_global, 1 subdivided)
int orc_numbering(int kind, int nc, int* node_dof) {
  return guarded([&] {
    Grid g;
    g.build(kind == 0 ? MeshKind::RefineGlobal : MeshKind::Subdivided, nc);
    std::memcpy(node_dof, g.node_dof.data(), sizeof(int) * g.n_dofs());
    return 0;
  });
}

// ---- MapMap (reference tests/test_tools.cpp:45-103) ---------------------------------------------
void* orc_mapmap_create() { return new MapMap(); }
void orc_mapmap_destroy(void* h) { delete static_cast<MapMap*>(h); }
void orc_mapmap_set(void* h, const double* px, int nx, const double* py, int ny, double det, const double* kkt3) {
  static_cast<MapMap*>(h)->set(std::vector<double>(px, px + nx), std::vector<double>(py, py + ny), det,
                               {kkt3[0], kkt3[1], kkt3[2]});
}
int orc_mapmap_get(void* h, const double* px, int nx, const double* py, int ny, double* det, double* kkt3) {
  return guarded([&] {
    std::array<double, 3> k;
    static_cast<MapMap*>(h)->get(std::vector<double>(px, px + nx), std::vector<double>(py, py + ny), *det, k);
    kkt3[0] = k[0];
    kkt3[1] = k[1];
    kkt3[2] = k[2];
    return 0;
  });
}
int orc_mapmap_size(void* h) { return static_cast<int>(static_cast<MapMap*>(h)->size()); }
int orc_mapmap_key(const double* px, int nx, const double* py, int ny, char* buf, int len) {
  std::string k = MapMap::to_string(std::vector<double>(px, px + nx), std::vector<double>(py, py + ny));
  std::strncpy(buf, k.c_str(), len - 1);
  buf[len - 1] = 0;
  return static_cast<int>(k.size());
}

// ---- elliptic-elliptic ---------------------------------------------------------------------
void* orc_ell_create(const char* prm, int macro_ref, int micro_ref) {
  try {
    auto* m = new elliptic::Manager(macro_ref, micro_ref, prm);
    m->setup();
    return m;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_ell_destroy(void* h) { delete static_cast<elliptic::Manager*>(h); }
void orc_ell_sizes(void* h, int* N, int* n) {
  auto* m = static_cast<elliptic::Manager*>(h);
  *N = m->micro_solver.num_grids;
  *n = m->micro_solver.grid.n_dofs();
}
void orc_ell_grid_locations(void* h, double* xy) {
  auto* m = static_cast<elliptic::Manager*>(h);
  for (int k = 0; k < m->micro_solver.num_grids; ++k) {
    xy[2 * k] = m->micro_solver.grid_locations[k][0];
    xy[2 * k + 1] = m->micro_solver.grid_locations[k][1];
  }
}
void orc_ell_dof_points(void* h, double* xy) {
  auto* m = static_cast<elliptic::Manager*>(h);
  for (int d = 0; d < m->micro_solver.grid.n_dofs(); ++d) m->micro_solver.grid.dof_point(d, xy + 2 * d);
}
// micro path for a given macro vector: assemble (+ solve)
int orc_ell_micro_run(void* h, const double* u, int assemble_only) {
  return guarded([&] {
    auto* m = static_cast<elliptic::Manager*>(h);
    std::copy(u, u + m->micro_solver.num_grids, m->macro_solver.solution.begin());
    m->micro_solver.assemble_system();
    if (!assemble_only) m->micro_solver.solve();
    return 0;
  });
}
void orc_ell_get_matrix_dense(void* h, int k, double* out) {
  dense_from(static_cast<elliptic::Manager*>(h)->micro_solver.system_matrices[k], out);
}
void orc_ell_get_rhs(void* h, int k, double* out) {
  auto& v = static_cast<elliptic::Manager*>(h)->micro_solver.righthandsides[k];
  std::copy(v.begin(), v.end(), out);
}
void orc_ell_get_solution(void* h, int k, double* out) {
  auto& v = static_cast<elliptic::Manager*>(h)->micro_solver.solutions[k];
  std::copy(v.begin(), v.end(), out);
}
void orc_ell_get_iters(void* h, int* out) {
  auto& v = static_cast<elliptic::Manager*>(h)->micro_solver.cg_iterations;
  std::copy(v.begin(), v.end(), out);
}
// recursive residuals of the last solve of grid k (entry i = after i steps); returns the count
int orc_ell_cg_history(void* h, int k, double* out, int maxlen) {
  auto& v = static_cast<elliptic::Manager*>(h)->micro_solver.cg_history[k];
  int n = std::min<int>(maxlen, static_cast<int>(v.size()));
  std::copy(v.begin(), v.begin() + n, out);
  return static_cast<int>(v.size());
}
void orc_ell_bulk(void* h, double* c) {
  auto* m = static_cast<elliptic::Manager*>(h);
  for (int k = 0; k < m->micro_solver.num_grids; ++k) c[k] = m->macro_solver.get_micro_bulk(k);
}
int orc_ell_set_exact_solution(void* h) {
  return guarded([&] {
    static_cast<elliptic::Manager*>(h)->micro_solver.set_exact_solution();
    return 0;
  });
}
void orc_ell_grid_errors(void* h, double* l2, double* h1) {
  auto* m = static_cast<elliptic::Manager*>(h);
  for (int k = 0; k < m->micro_solver.num_grids; ++k) m->micro_solver.grid_errors(k, l2[k], h1[k]);
}
// full driver
int orc_ell_run(void* h, int cycle_limit) {
  return guarded([&] {
    auto* m = static_cast<elliptic::Manager*>(h);
    m->cycle_limit = cycle_limit;
    m->run();
    return static_cast<int>(m->history.size());
  });
}
int orc_ell_macro_dofs(void* h) { return static_cast<elliptic::Manager*>(h)->macro_solver.grid.n_dofs(); }
void orc_ell_history(void* h, int cycle, double* errs4, double* macro_sol, double* contrib, double* micro_sols,
                     int* micro_its, int* macro_its) {
  auto* m = static_cast<elliptic::Manager*>(h);
  const auto& r = m->history.at(cycle);
  errs4[0] = r.micro_l2;
  errs4[1] = r.micro_h1;
  errs4[2] = r.macro_l2;
  errs4[3] = r.macro_h1;
  if (macro_sol) std::copy(r.macro_solution.begin(), r.macro_solution.end(), macro_sol);
  if (contrib) std::copy(r.micro_contribution.begin(), r.micro_contribution.end(), contrib);
  if (micro_sols) {
    size_t n = m->micro_solver.grid.n_dofs();
    for (size_t k = 0; k < r.micro_solutions.size(); ++k)
      std::copy(r.micro_solutions[k].begin(), r.micro_solutions[k].end(), micro_sols + k * n);
  }
  if (micro_its) std::copy(r.micro_cg_iterations.begin(), r.micro_cg_iterations.end(), micro_its);
  if (macro_its) *macro_its = r.macro_cg_iterations;
}
int orc_ell_table(void* h, char* buf, int len) {
  std::string t = static_cast<elliptic::Manager*>(h)->convergence_table();
  std::strncpy(buf, t.c_str(), len - 1);
  buf[len - 1] = 0;
  return static_cast<int>(t.size());
}

// Timed micro path (CPU baseline): assemble+solve of systems [k0,k1) with `threads` std::threads
// over k, the way solve_biomath spreads its grids over TBB tasks.  Returns seconds.
double orc_ell_time_micro(void* h, const double* u, int k0, int k1, int threads) {
  auto* m = static_cast<elliptic::Manager*>(h);
  std::copy(u, u + m->micro_solver.num_grids, m->macro_solver.solution.begin());
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([=] {
      for (int k = k0 + t; k < k1; k += threads) {
        m->micro_solver.assemble_one(k);
        try {
          m->micro_solver.solve_one(k);
        } catch (const std::exception&) {
          // SolverControl::NoConvergence after 1000 steps (the reference would stop here): the time is what is
          // being measured, the iterate after 1000 steps stays in place
        }
      }
    });
  for (auto& th : pool) th.join();
  for (int k = k0; k < k1; ++k) (void)m->macro_solver.get_micro_bulk(k);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// The same micro path with the reference's string-keyed MapMap in the inner loops - "what the original pays"
// (SURVEY 8d): compute_pullback_objects fills the two std::map<std::string, ...> (micro.cpp:75-94, mapping.cpp:15-27),
// integrate_cell looks (det_jac, kkt) up per quadrature point and det_jac again per (test function, point)
// (micro.cpp:108-136), get_micro_bulk once more per point (macro.cpp:134-157); every lookup formats its key with an
// ostringstream (mapping.cpp:47-57).  One thread, as solve_elliptic runs.  Returns the seconds of assemble + solve +
// bulk integral for the systems [k0, k1); *build_seconds = the one-off cost of filling the map for them.
double orc_ell_time_micro_mapmap(void* h, const double* u, int k0, int k1, double* build_seconds) {
  auto* m = static_cast<elliptic::Manager*>(h);
  auto& ms = m->micro_solver;
  std::copy(u, u + ms.num_grids, m->macro_solver.solution.begin());
  const int p = ms.fem_q_deg, nq = p * p;
  const double hh = ms.grid.h;
  MapMap mapmap;
  auto tb = std::chrono::steady_clock::now();
  for (int c = 0; c < ms.grid.n_cells(); ++c)
    for (int k = k0; k < k1; ++k)
      for (int q = 0; q < nq; ++q) {
        double y[2];
        ms.qpoint(c, q, y);
        const size_t t = ms.table_index(k, c, q);
        mapmap.set({ms.grid_locations[k][0], ms.grid_locations[k][1]}, {y[0], y[1]}, ms.det_jac[t],
                   {ms.kkt[3 * t], ms.kkt[3 * t + 1], ms.kkt[3 * t + 2]});
      }
  if (build_seconds) *build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count();
  auto t0 = std::chrono::steady_clock::now();
  for (int k = k0; k < k1; ++k) {
    const std::vector<double> px = {ms.grid_locations[k][0], ms.grid_locations[k][1]};
    std::fill(ms.righthandsides[k].begin(), ms.righthandsides[k].end(), 0.0);
    std::fill(ms.solutions[k].begin(), ms.solutions[k].end(), 0.0);
    ms.system_matrices[k].reinit(ms.pattern);
    for (int c = 0; c < ms.grid.n_cells(); ++c) {
      double cell_matrix[4][4] = {};
      double cell_rhs[4] = {};
      for (int q = 0; q < nq; ++q) {
        double y[2], dj;
        std::array<double, 3> kk;
        ms.qpoint(c, q, y);
        mapmap.get(px, {y[0], y[1]}, dj, kk);
        const int qi = q % p, qj = q / p;
        double g[4][2];
        q1_grads(ms.qx[qi], ms.qx[qj], g);
        const double JxW = ms.qw[qi] * ms.qw[qj] * hh * hh;
        for (int i = 0; i < 4; ++i) {
          const double gi0 = g[i][0] / hh, gi1 = g[i][1] / hh;
          const double t0_ = gi0 * kk[0] + gi1 * kk[1], t1_ = gi0 * kk[1] + gi1 * kk[2];
          for (int j = 0; j < 4; ++j) cell_matrix[i][j] += (t0_ * (g[j][0] / hh) + t1_ * (g[j][1] / hh)) * JxW * dj;
        }
      }
      for (int i = 0; i < 4; ++i)
        for (int q = 0; q < nq; ++q) {
          double y[2], mapped[2], v[4], dj;
          ms.qpoint(c, q, y);
          mapmap.get_det_jac(px, {y[0], y[1]}, dj);
          const int qi = q % p, qj = q / p;
          ms.data.mapping.mmap(ms.grid_locations[k].data(), y, mapped);
          const double rhs_val = ms.data.micro_rhs.mvalue(ms.grid_locations[k].data(), mapped);
          q1_values(ms.qx[qi], ms.qx[qj], v);
          cell_rhs[i] += (u[k] + rhs_val) * v[i] * (ms.qw[qi] * ms.qw[qj] * hh * hh) * dj;
        }
      const auto& cd = ms.grid.cell_dofs[c];
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) ms.system_matrices[k].add(cd[i], cd[j], cell_matrix[i][j]);
        ms.righthandsides[k][cd[i]] += cell_rhs[i];
      }
    }
    std::map<int, double> bv;
    for (int node = 0; node < ms.grid.n_dofs(); ++node)
      if (ms.grid.node_on_boundary(node)) {
        double y[2] = {-1.0 + (node % ms.grid.m) * ms.grid.h, -1.0 + (node / ms.grid.m) * ms.grid.h};
        bv[ms.grid.node_dof[node]] = ms.data.micro_bc.mvalue(ms.grid_locations[k].data(), y);
      }
    apply_boundary_values(bv, ms.system_matrices[k], ms.solutions[k], ms.righthandsides[k]);
    const CGResult r = solve_cg(ms.system_matrices[k], ms.solutions[k], ms.righthandsides[k], 1000, 1e-12, nullptr);
    ms.cg_iterations[k] = r.iterations;
    double integral = 0;
    for (int c = 0; c < ms.grid.n_cells(); ++c)
      for (int q = 0; q < nq; ++q) {
        double y[2], dj;
        ms.qpoint(c, q, y);
        mapmap.get_det_jac(px, {y[0], y[1]}, dj);
        const int qi = q % p, qj = q / p;
        integral += q1_interp(ms.grid, ms.solutions[k], c, ms.qx[qi], ms.qx[qj]) * dj * (ms.qw[qi] * ms.qw[qj] * hh * hh);
      }
    m->macro_solver.micro_contribution[k] = integral;
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"

#include "capi_bio.inc"
#include "capi_parabolic.inc"
