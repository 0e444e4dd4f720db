// Host-side building blocks of the (small) macro problems: Q1 on a square mesh of [-1,1]^2 in the
// reference's DoF numbering, CSR matrix, Dirichlet elimination, CG, nodal-function norms, and
// parsed macro functions.  The macro systems have N unknowns with a 9-point stencil and are solved
// once per fixed-point iteration (reference: src/elliptic-elliptic/macro.cpp:57-114); SURVEY.md
// §8f ranks moving them onto GPU 0 as the next step after the micro path.
//
// Macro functions of (x0,x1[,t]) are compiled with the same expression compiler as the micro
// functions and evaluated by running the VM on the host (expr_vm.cuh is host/device code).
#pragma once
#include <thread>
#include <cstdlib>
#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../csrc/expr_compile.h"
#include "../csrc/expr_vm.cuh"
#include "../csrc/fe_tables.h"

namespace ms {

// A parsed function of the macro point (and time): FunctionParser<2> of the reference.
class MacroFunction {
 public:
  void initialize(const std::string& expression, const std::map<std::string, double>& constants,
                  bool time_dependent = false) {
    set_ = msgpu::ProgramSet();
    const int node = set_.parse_function(expression, constants, time_dependent, 1)[0];
    prog_ = set_.add_point_program({node}, msgpu::MODE_POINTS);
    set_.finalize();
    slots_.assign(std::max(1, set_.n_slots()), 0.0);
  }
  void set_time(double t) { time_ = t; }
  double value(double x0, double x1) const {  // thread-safe: all scratch is local
    double regs[16] = {x0, x1, time_};
    double small[32];
    std::vector<double> big;
    double* slots = small;
    if (slots_.size() > 32) {
      big.assign(slots_.size(), 0.0);
      slots = big.data();
    }
    msgpu::VmIo io{};
    io.R = regs;
    io.r_stride = 1;
    io.OUT = slots;
    io.U = slots;
    io.out_stride = 1;
    if (set_.n_slots() > 0) msgpu::vm_run(set_.slot_program(), 0, io);
    const msgpu::VmProgram& p = set_.point_program(prog_);
    if (p.n_regs > 16) throw std::runtime_error("macro expression needs too many registers");
    msgpu::vm_run(p, 0, io);
    return regs[p.out0];
  }
  // AutoDerivativeFunction gradient: central differences, h = 1e-8
  void gradient(double x0, double x1, double g[2]) const {
    const double h = 1e-8;
    g[0] = (value(x0 + h, x1) - value(x0 - h, x1)) / (2 * h);
    g[1] = (value(x0, x1 + h) - value(x0, x1 - h)) / (2 * h);
  }

 private:
  msgpu::ProgramSet set_;
  int prog_ = -1;
  double time_ = 0;
  mutable std::vector<double> slots_;
};

// Square mesh with nc x nc cells; DoFs in the reference numbering (see fe_tables.h).
struct MacroMesh {
  int nc = 0, m = 0, n = 0;
  double h = 0;
  std::vector<int> dof_ix, dof_iy;              // lattice position of every DoF
  std::vector<std::array<int, 4>> cell_dofs;    // active-cell order of the reference
  std::vector<int> cell_cx, cell_cy;
  std::vector<int> rowstart, col;               // CSR pattern, diagonal first then ascending

  void build(int mesh_kind, int cells_per_axis) {
    nc = cells_per_axis;
    m = nc + 1;
    n = m * m;
    h = 2.0 / nc;
    const std::vector<int> ref_to_internal = msgpu::reference_numbering(mesh_kind, nc);
    std::vector<int> internal_to_ref(n);
    dof_ix.resize(n);
    dof_iy.resize(n);
    for (int i = 0; i < n; ++i) {
      internal_to_ref[ref_to_internal[i]] = i;
      dof_ix[i] = ref_to_internal[i] / m;
      dof_iy[i] = ref_to_internal[i] % m;
    }
    int levels = 0;
    while ((1 << levels) < nc) ++levels;
    cell_dofs.resize(nc * nc);
    cell_cx.resize(nc * nc);
    cell_cy.resize(nc * nc);
    for (int a = 0; a < nc * nc; ++a) {
      int cx = 0, cy = 0;
      if (mesh_kind == 0) {
        for (int b = 0; b < levels; ++b) {
          cx |= ((a >> (2 * b)) & 1) << b;
          cy |= ((a >> (2 * b + 1)) & 1) << b;
        }
      } else {
        cx = a % nc;
        cy = a / nc;
      }
      cell_cx[a] = cx;
      cell_cy[a] = cy;
      for (int v = 0; v < 4; ++v) cell_dofs[a][v] = internal_to_ref[(cx + (v & 1)) * m + (cy + (v >> 1))];
    }
    std::vector<std::vector<int>> rows(n);
    for (const auto& cd : cell_dofs)
      for (int a : cd)
        for (int b : cd) rows[a].push_back(b);
    rowstart.assign(n + 1, 0);
    col.clear();
    for (int r = 0; r < n; ++r) {
      std::sort(rows[r].begin(), rows[r].end());
      rows[r].erase(std::unique(rows[r].begin(), rows[r].end()), rows[r].end());
      col.push_back(r);
      for (int c : rows[r])
        if (c != r) col.push_back(c);
      rowstart[r + 1] = static_cast<int>(col.size());
    }
  }
  double x_of(int dof) const { return -1.0 + dof_ix[dof] * h; }
  double y_of(int dof) const { return -1.0 + dof_iy[dof] * h; }
  bool on_boundary(int dof) const {
    return dof_ix[dof] == 0 || dof_iy[dof] == 0 || dof_ix[dof] == m - 1 || dof_iy[dof] == m - 1;
  }
  int entry(int r, int c) const {
    if (col[rowstart[r]] == c) return rowstart[r];
    auto b = col.begin() + rowstart[r] + 1, e = col.begin() + rowstart[r + 1];
    auto it = std::lower_bound(b, e, c);
    if (it == e || *it != c) throw std::runtime_error("entry outside the sparsity pattern");
    return static_cast<int>(it - col.begin());
  }
};

struct CsrMatrix {
  const MacroMesh* mesh = nullptr;
  std::vector<double> val;
  void reinit(const MacroMesh& ms_) {
    mesh = &ms_;
    val.assign(ms_.col.size(), 0.0);
  }
  void zero() { std::fill(val.begin(), val.end(), 0.0); }
  void add(int r, int c, double v) { val[mesh->entry(r, c)] += v; }
  double diag(int r) const { return val[mesh->rowstart[r]]; }
  void vmult(std::vector<double>& y, const std::vector<double>& x) const {
    for (int r = 0; r < mesh->n; ++r) {
      double s = 0;
      for (int k = mesh->rowstart[r]; k < mesh->rowstart[r + 1]; ++k) s += val[k] * x[mesh->col[k]];
      y[r] = s;
    }
  }
};

// MatrixTools::apply_boundary_values with column elimination (macro.cpp:105).
inline void apply_boundary_values(const std::map<int, double>& values, CsrMatrix& A, std::vector<double>& x,
                                  std::vector<double>& b) {
  const MacroMesh& ms_ = *A.mesh;
  double first_diag = 1.0;
  for (int i = 0; i < ms_.n; ++i)
    if (A.diag(i) != 0.0) {
      first_diag = A.diag(i);
      break;
    }
  for (const auto& bv : values) {
    const int d = bv.first;
    for (int k = ms_.rowstart[d] + 1; k < ms_.rowstart[d + 1]; ++k) A.val[k] = 0.0;
    if (A.diag(d) == 0.0) A.val[ms_.rowstart[d]] = first_diag;
    const double a_dd = A.diag(d), new_rhs = bv.second * a_dd;
    b[d] = new_rhs;
    for (int k = ms_.rowstart[d] + 1; k < ms_.rowstart[d + 1]; ++k) {
      const int row = ms_.col[k], p = ms_.entry(row, d);
      b[row] -= A.val[p] / a_dd * new_rhs;
      A.val[p] = 0.0;
    }
    x[d] = bv.second;
  }
}

// SolverCG + SolverControl(max_steps, tol) + PreconditionIdentity; returns the step count,
// throws where deal.II throws SolverControl::NoConvergence.
inline int solve_cg(const CsrMatrix& A, std::vector<double>& x, const std::vector<double>& b, int max_steps,
                    double tol) {
  const int n = A.mesh->n;
  std::vector<double> g(n), d(n), h(n);
  auto dot = [n](const std::vector<double>& p, const std::vector<double>& q) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += p[i] * q[i];
    return s;
  };
  A.vmult(g, x);
  for (int i = 0; i < n; ++i) g[i] -= b[i];
  double res = std::sqrt(dot(g, g));
  if (res <= tol) return 0;
  for (int i = 0; i < n; ++i) d[i] = -g[i];
  double gh = res * res;
  for (int it = 1;; ++it) {
    A.vmult(h, d);
    const double alpha = gh / dot(d, h);
    for (int i = 0; i < n; ++i) {
      x[i] += alpha * d[i];
      g[i] += alpha * h[i];
    }
    res = std::sqrt(dot(g, g));
    if (res <= tol) return it;
    if (it >= max_steps || res != res) throw std::runtime_error("SolverControl::NoConvergence (macro system)");
    const double beta = res * res / gh;
    gh = res * res;
    for (int i = 0; i < n; ++i) d[i] = beta * d[i] - g[i];
  }
}

struct Quadrature {
  std::vector<double> pt, wt;
  explicit Quadrature(int q) { msgpu::gauss_legendre_unit(q, pt, wt); }
  int size() const { return static_cast<int>(pt.size()); }
};

inline void shape(double xi, double eta, double v[4]) {
  v[0] = (1 - xi) * (1 - eta);
  v[1] = xi * (1 - eta);
  v[2] = (1 - xi) * eta;
  v[3] = xi * eta;
}
inline void shape_grad(double xi, double eta, double h, double g[4][2]) {
  g[0][0] = -(1 - eta) / h; g[0][1] = -(1 - xi) / h;
  g[1][0] = (1 - eta) / h;  g[1][1] = -xi / h;
  g[2][0] = -eta / h;       g[2][1] = (1 - xi) / h;
  g[3][0] = eta / h;        g[3][1] = xi / h;
}

// sign * int phi_a phi_b with QGauss(q) on every cell: the operator C of  b(c) = b0 + C c  when the
// interpolated micro contribution enters a macro load linearly (macro.cpp:78-93, bio macro.cpp:144-165).
inline CsrMatrix mass_operator(const MacroMesh& mesh, int q, double sign) {
  CsrMatrix C;
  C.reinit(mesh);
  Quadrature Q(q);
  const double h = mesh.h;
  double M[4][4] = {};
  for (int j = 0; j < q; ++j)
    for (int i = 0; i < q; ++i) {
      double v[4];
      shape(Q.pt[i], Q.pt[j], v);
      const double jxw = Q.wt[i] * Q.wt[j] * h * h;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) M[a][b] += sign * v[a] * v[b] * jxw;
    }
  for (const auto& cd : mesh.cell_dofs)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) C.add(cd[a], cd[b], M[a][b]);
  return C;
}
inline void zero_rows(CsrMatrix& C, const std::map<int, double>& rows) {
  for (const auto& r : rows)
    for (int k = C.mesh->rowstart[r.first]; k < C.mesh->rowstart[r.first + 1]; ++k) C.val[k] = 0.0;
}
// MSGPU_HOST_MACRO=1 keeps the macro systems on the host (A/B against the device path)
inline bool device_macro_enabled() {
  const char* e = std::getenv("MSGPU_HOST_MACRO");
  return !(e && e[0] == '1');
}

// || sum_k values[k] phi_k ||_{L2} with QGauss(q): integrate_difference(.., ZeroFunction, L2_norm)
// followed by compute_global_error (micro.cpp:206-214).
inline double nodal_l2_norm(const MacroMesh& mesh, const std::vector<double>& values, int q) {
  Quadrature Q(q);
  // basis values and weights of the q x q points once (they are the same in every cell); the per-cell
  // arithmetic and its order are unchanged, cells are spread over host threads and the per-cell results are
  // added in cell order afterwards, so the value does not depend on the thread count
  std::vector<std::array<double, 4>> phi(static_cast<size_t>(q) * q);
  std::vector<double> wi(static_cast<size_t>(q) * q), wj(static_cast<size_t>(q) * q);
  for (int j = 0; j < q; ++j)
    for (int i = 0; i < q; ++i) {
      shape(Q.pt[i], Q.pt[j], phi[static_cast<size_t>(j) * q + i].data());
      wi[static_cast<size_t>(j) * q + i] = Q.wt[i];
      wj[static_cast<size_t>(j) * q + i] = Q.wt[j];
    }
  const double h = mesh.h;
  const size_t n_cells = mesh.cell_dofs.size(), nq = phi.size();
  std::vector<double> cell_sq(n_cells);
  auto work = [&](size_t c0, size_t c1) {
    for (size_t c = c0; c < c1; ++c) {
      const auto& cd = mesh.cell_dofs[c];
      const double u0 = values[cd[0]], u1 = values[cd[1]], u2 = values[cd[2]], u3 = values[cd[3]];
      double s = 0;
      for (size_t p = 0; p < nq; ++p) {
        const auto& v = phi[p];
        double f = 0;
        f += v[0] * u0;
        f += v[1] * u1;
        f += v[2] * u2;
        f += v[3] * u3;
        s += f * f * wi[p] * wj[p] * h * h;  // same association as the reference-order loop it replaces
      }
      const double cell = std::sqrt(s);
      cell_sq[c] = cell * cell;
    }
  };
  unsigned int T = std::thread::hardware_concurrency();
  if (T > 32) T = 32;
  if (T < 2 || n_cells < 1024) {
    work(0, n_cells);
  } else {
    std::vector<std::thread> pool;
    const size_t chunk = (n_cells + T - 1) / T;
    for (unsigned int t = 0; t < T; ++t) {
      const size_t c0 = t * chunk, c1 = std::min(n_cells, c0 + chunk);
      if (c0 < c1) pool.emplace_back(work, c0, c1);
    }
    for (auto& th : pool) th.join();
  }
  double total = 0;
  for (size_t c = 0; c < n_cells; ++c) total += cell_sq[c];
  return std::sqrt(total);
}

// L2 and H1-seminorm errors of a nodal field against a parsed function, QGauss(q).
inline void field_errors(const MacroMesh& mesh, const std::vector<double>& u, const MacroFunction& exact, int q,
                         double& l2, double& h1) {
  Quadrature Q(q);
  // cells are independent (5 expression evaluations per point): host threads over cell ranges, per-cell
  // sums added in cell order afterwards, so the result does not depend on the thread count
  const size_t n_cells = mesh.cell_dofs.size();
  std::vector<double> cell_l2(n_cells), cell_h1(n_cells);
  auto work = [&](size_t c0, size_t c1) {
  for (size_t c = c0; c < c1; ++c) {
    const double ox = -1.0 + mesh.cell_cx[c] * mesh.h, oy = -1.0 + mesh.cell_cy[c] * mesh.h;
    double cl2 = 0, ch1 = 0;
    for (int j = 0; j < q; ++j)
      for (int i = 0; i < q; ++i) {
        double v[4], g[4][2];
        shape(Q.pt[i], Q.pt[j], v);
        shape_grad(Q.pt[i], Q.pt[j], mesh.h, g);
        const double x0 = ox + mesh.h * Q.pt[i], x1 = oy + mesh.h * Q.pt[j];
        const double jxw = Q.wt[i] * Q.wt[j] * mesh.h * mesh.h;
        double uh = 0, gx = 0, gy = 0;
        for (int a = 0; a < 4; ++a) {
          const double ua = u[mesh.cell_dofs[c][a]];
          uh += ua * v[a];
          gx += ua * g[a][0];
          gy += ua * g[a][1];
        }
        double eg[2];
        exact.gradient(x0, x1, eg);
        const double d = uh - exact.value(x0, x1);
        cl2 += d * d * jxw;
        ch1 += ((gx - eg[0]) * (gx - eg[0]) + (gy - eg[1]) * (gy - eg[1])) * jxw;
      }
    cell_l2[c] = cl2;
    cell_h1[c] = ch1;
  }
  };
  unsigned int T = std::thread::hardware_concurrency();
  if (T > 32) T = 32;
  if (T < 2 || n_cells < 256) {
    work(0, n_cells);
  } else {
    std::vector<std::thread> pool;
    const size_t chunk = (n_cells + T - 1) / T;
    for (unsigned int t = 0; t < T; ++t) {
      const size_t c0 = t * chunk, c1 = std::min(n_cells, c0 + chunk);
      if (c0 < c1) pool.emplace_back(work, c0, c1);
    }
    for (auto& th : pool) th.join();
  }
  double sl2 = 0, sh1 = 0;
  for (size_t c = 0; c < n_cells; ++c) {
    sl2 += cell_l2[c];
    sh1 += cell_h1[c];
  }
  l2 = std::sqrt(sl2);
  h1 = std::sqrt(sh1);
}

}  // namespace ms
