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
        m->micro_solver.solve_one(k);
      }
    });
  for (auto& th : pool) th.join();
  for (int k = k0; k < k1; ++k) (void)m->macro_solver.get_micro_bulk(k);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"

#include "capi_bio.inc"
#include "capi_parabolic.inc"
