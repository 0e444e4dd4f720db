// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// The deal.II 9.1.1 building blocks the reference's micro/macro solvers call,
// restated for the only case the reference uses: FE_Q<2>(1) on axis-aligned
// square cells of [-1,1]^2.  deal.II is a third-party dependency that is not
// vendored under /root/reference (CMakeLists.txt:10,27), so what follows is a
// restatement of its documented behaviour at the reference's call sites
// ("parity unpinned": no golden vectors exist in the reference for this part).
//
//   QGauss<1>(p)                      -> gauss01()
//   GridGenerator::hyper_cube+refine_global / subdivided_hyper_cube,
//   DoFHandler::distribute_dofs(FE_Q(1)) -> Grid (cell order + first-touch numbering)
//   DoFTools::make_sparsity_pattern   -> Grid::build_pattern (diagonal first, then ascending)
//   SparseMatrix<double>              -> SparseMatrix
//   MatrixTools::apply_boundary_values-> apply_boundary_values
//   SolverCG<>/SolverControl/PreconditionIdentity -> solve_cg
//   FEValues / FEFaceValues           -> Q1 helpers
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <vector>

namespace orc {

// p-point Gauss-Legendre rule on [0,1], ascending nodes.
inline void gauss01(int p, std::vector<double>& x, std::vector<double>& w) {
  x.assign(p, 0.0);
  w.assign(p, 0.0);
  const long double pi = 3.141592653589793238462643383279502884L;
  for (int i = 0; i < (p + 1) / 2; ++i) {
    long double z = std::cos(pi * (i + 0.75L) / (p + 0.5L));
    long double pp = 0;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1, p2 = 0;
      for (int j = 0; j < p; ++j) {
        long double p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j + 1.0L) * z * p2 - j * p3) / (j + 1.0L);
      }
      pp = p * (z * p1 - p2) / (z * z - 1.0L);
      long double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs((double)(z - z1)) < 1e-19) break;
    }
    long double xi = 0.5L * (1.0L - z), xj = 0.5L * (1.0L + z);
    long double wi = 1.0L / ((1.0L - z * z) * pp * pp);
    x[i] = (double)xi;
    x[p - 1 - i] = (double)xj;
    w[i] = (double)wi;
    w[p - 1 - i] = (double)wi;
  }
}

// Q1 shape functions on the unit cell, deal.II vertex order (0,0),(1,0),(0,1),(1,1).
inline void q1_values(double xi, double eta, double v[4]) {
  v[0] = (1 - xi) * (1 - eta);
  v[1] = xi * (1 - eta);
  v[2] = (1 - xi) * eta;
  v[3] = xi * eta;
}
inline void q1_grads(double xi, double eta, double g[4][2]) {
  g[0][0] = -(1 - eta); g[0][1] = -(1 - xi);
  g[1][0] = (1 - eta);  g[1][1] = -xi;
  g[2][0] = -eta;       g[2][1] = (1 - xi);
  g[3][0] = eta;        g[3][1] = xi;
}

enum class MeshKind { RefineGlobal = 0, Subdivided = 1 };

struct SparsityPattern {
  int n = 0;
  std::vector<int> rowstart;  // n+1
  std::vector<int> col;       // diagonal first in each row, then ascending
  int find(int r, int c) const {
    if (col[rowstart[r]] == c) return rowstart[r];
    auto b = col.begin() + rowstart[r] + 1, e = col.begin() + rowstart[r + 1];
    auto it = std::lower_bound(b, e, c);
    if (it == e || *it != c) return -1;
    return static_cast<int>(it - col.begin());
  }
};

// A square mesh of [-1,1]^2 with nc x nc cells, carrying deal.II's active-cell
// order and Q1 DoF numbering.
struct Grid {
  int nc = 0;          // cells per axis
  int m = 0;           // nodes per axis
  double h = 0;        // cell side
  MeshKind kind = MeshKind::Subdivided;
  std::vector<int> cell_ix, cell_iy;     // active-cell order -> lattice position
  std::vector<int> node_dof;             // lexicographic node (iy*m+ix) -> dof index
  std::vector<int> dof_node;             // inverse
  std::vector<std::array<int, 4>> cell_dofs;

  void build(MeshKind k, int cells_per_axis) {
    kind = k;
    nc = cells_per_axis;
    m = nc + 1;
    h = 2.0 / nc;
    const int ncell = nc * nc;
    cell_ix.resize(ncell);
    cell_iy.resize(ncell);
    if (k == MeshKind::Subdivided) {
      for (int c = 0; c < ncell; ++c) {
        cell_ix[c] = c % nc;
        cell_iy[c] = c / nc;
      }
    } else {
      int r = 0;
      while ((1 << r) < nc) ++r;
      if ((1 << r) != nc) throw std::runtime_error("refine_global mesh needs 2^r cells per axis");
      for (int c = 0; c < ncell; ++c) {
        int ix = 0, iy = 0;
        for (int l = r - 1; l >= 0; --l) {
          int child = (c >> (2 * l)) & 3;
          ix = ix * 2 + (child & 1);
          iy = iy * 2 + (child >> 1);
        }
        cell_ix[c] = ix;
        cell_iy[c] = iy;
      }
    }
    node_dof.assign(m * m, -1);
    int next = 0;
    cell_dofs.resize(ncell);
    for (int c = 0; c < ncell; ++c) {
      for (int v = 0; v < 4; ++v) {
        int nx = cell_ix[c] + (v & 1), ny = cell_iy[c] + (v >> 1);
        int& d = node_dof[ny * m + nx];
        if (d < 0) d = next++;
        cell_dofs[c][v] = d;
      }
    }
    dof_node.assign(m * m, -1);
    for (int i = 0; i < m * m; ++i) dof_node[node_dof[i]] = i;
  }
  int n_dofs() const { return m * m; }
  int n_cells() const { return nc * nc; }
  double cell_x0(int c) const { return -1.0 + cell_ix[c] * h; }
  double cell_y0(int c) const { return -1.0 + cell_iy[c] * h; }
  void dof_point(int dof, double p[2]) const {
    int node = dof_node[dof];
    p[0] = -1.0 + (node % m) * h;
    p[1] = -1.0 + (node / m) * h;
  }
  bool node_on_boundary(int node) const {
    int ix = node % m, iy = node / m;
    return ix == 0 || iy == 0 || ix == m - 1 || iy == m - 1;
  }
  // Is face f (0:x-min 1:x-max 2:y-min 3:y-max) of cell c on the boundary?
  bool face_at_boundary(int c, int f) const {
    switch (f) {
      case 0: return cell_ix[c] == 0;
      case 1: return cell_ix[c] == nc - 1;
      case 2: return cell_iy[c] == 0;
      default: return cell_iy[c] == nc - 1;
    }
  }
  void build_pattern(SparsityPattern& sp) const {
    const int n = n_dofs();
    std::vector<std::vector<int>> rows(n);
    for (const auto& cd : cell_dofs)
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) rows[cd[a]].push_back(cd[b]);
    sp.n = n;
    sp.rowstart.assign(n + 1, 0);
    sp.col.clear();
    for (int r = 0; r < n; ++r) {
      auto& v = rows[r];
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      sp.col.push_back(r);
      for (int c : v)
        if (c != r) sp.col.push_back(c);
      sp.rowstart[r + 1] = static_cast<int>(sp.col.size());
    }
  }
};

struct SparseMatrix {
  const SparsityPattern* sp = nullptr;
  std::vector<double> val;
  void reinit(const SparsityPattern& p) {
    sp = &p;
    val.assign(p.col.size(), 0.0);
  }
  void add(int r, int c, double v) {
    int k = sp->find(r, c);
    if (k < 0) throw std::runtime_error("SparseMatrix::add outside pattern");
    val[k] += v;
  }
  double diag(int r) const { return val[sp->rowstart[r]]; }
  void vmult(std::vector<double>& dst, const std::vector<double>& src) const {
    const int n = sp->n;
    for (int r = 0; r < n; ++r) {
      double s = 0;
      for (int k = sp->rowstart[r]; k < sp->rowstart[r + 1]; ++k) s += val[k] * src[sp->col[k]];
      dst[r] = s;
    }
  }
  void copy_from(const SparseMatrix& o) {
    sp = o.sp;
    val = o.val;
  }
  void add_scaled(double f, const SparseMatrix& o) {
    for (size_t i = 0; i < val.size(); ++i) val[i] += f * o.val[i];
  }
};

// MatrixTools::apply_boundary_values(boundary_values, A, x, b, eliminate_columns = true)
// as called at reference src/elliptic-elliptic/micro.cpp:168, macro.cpp:105.
inline void apply_boundary_values(const std::map<int, double>& bv, SparseMatrix& A, std::vector<double>& x,
                                  std::vector<double>& b) {
  if (bv.empty()) return;
  const SparsityPattern& sp = *A.sp;
  double first_nonzero_diag = 1;
  for (int i = 0; i < sp.n; ++i)
    if (A.diag(i) != 0) {
      first_nonzero_diag = A.diag(i);
      break;
    }
  for (const auto& kv : bv) {
    const int d = kv.first;
    const double v = kv.second;
    for (int k = sp.rowstart[d] + 1; k < sp.rowstart[d + 1]; ++k) A.val[k] = 0.0;
    double new_rhs;
    if (A.diag(d) != 0) {
      new_rhs = v * A.diag(d);
    } else {
      A.val[sp.rowstart[d]] = first_nonzero_diag;
      new_rhs = v * first_nonzero_diag;
    }
    b[d] = new_rhs;
    const double diagonal_entry = A.diag(d);
    for (int k = sp.rowstart[d] + 1; k < sp.rowstart[d + 1]; ++k) {
      const int row = sp.col[k];
      const int p = sp.find(row, d);
      b[row] -= A.val[p] / diagonal_entry * new_rhs;
      A.val[p] = 0.0;
    }
    x[d] = v;
  }
}

struct CGResult {
  int iterations = 0;
  double residual = 0;
  bool converged = false;
};

// deal.II 9.1 SolverCG<Vector<double>>::solve with PreconditionIdentity and
// SolverControl(max_steps, tol): recursive residual, absolute tolerance.
// Called at reference micro.cpp:176-179, bio micro.cpp:310-313, rho_solver.cpp:223-226.
inline CGResult solve_cg(const SparseMatrix& A, std::vector<double>& x, const std::vector<double>& b,
                         int max_steps, double tol, std::vector<double>* history = nullptr) {
  const int n = A.sp->n;
  std::vector<double> g(n), d(n), h(n);
  bool all_zero = true;
  for (double v : x)
    if (v != 0) {
      all_zero = false;
      break;
    }
  if (!all_zero) {
    A.vmult(g, x);
    for (int i = 0; i < n; ++i) g[i] -= b[i];
  } else {
    for (int i = 0; i < n; ++i) g[i] = -b[i];
  }
  auto dot = [n](const std::vector<double>& a, const std::vector<double>& c) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += a[i] * c[i];
    return s;
  };
  CGResult r;
  double res = std::sqrt(dot(g, g));
  r.residual = res;
  if (history) {
    history->clear();
    history->push_back(res);
  }
  if (res <= tol) {
    r.converged = true;
    return r;
  }
  if (max_steps <= 0 || std::isnan(res)) return r;
  for (int i = 0; i < n; ++i) d[i] = -g[i];
  double gh = res * res;
  int it = 0;
  for (;;) {
    ++it;
    A.vmult(h, d);
    double alpha = dot(d, h);
    alpha = gh / alpha;
    for (int i = 0; i < n; ++i) x[i] += alpha * d[i];
    for (int i = 0; i < n; ++i) g[i] += alpha * h[i];
    res = std::sqrt(dot(g, g));
    r.iterations = it;
    r.residual = res;
    if (history) history->push_back(res);
    if (res <= tol) {
      r.converged = true;
      return r;
    }
    if (it >= max_steps || std::isnan(res)) return r;
    double beta = gh;
    gh = res * res;
    beta = gh / beta;
    for (int i = 0; i < n; ++i) d[i] = beta * d[i] - g[i];
  }
}

// VectorTools::project(mapping, dof_handler, constraints, QGauss<2>(p), f, vec) for Q1 on the square mesh:
// mass matrix and load vector with the given rule (the Q1 mass matrix is integrated exactly for p >= 2),
// CG to 1e-12 relative to the load (deal.II 9.1: ReductionControl(5 n, 0., 1e-12)).  Called by the
// reference at src/elliptic-elliptic/micro.cpp:231, src/bio-multiscale/micro.cpp:388,
// src/elliptic-parabolic/rho_solver.cpp:357.  f(y) = value at the micro point y.
template <class F>
inline std::vector<double> l2_project(const Grid& grid, const SparsityPattern& pattern, int p, F&& f) {
  std::vector<double> qx, qw;
  gauss01(p, qx, qw);
  const double h = grid.h;
  const int n = grid.n_dofs();
  SparseMatrix mass;
  mass.reinit(pattern);
  std::vector<double> rhs(n, 0.0), sol(n, 0.0);
  for (int c = 0; c < grid.n_cells(); ++c) {
    const auto& cd = grid.cell_dofs[c];
    for (int qj = 0; qj < p; ++qj)
      for (int qi = 0; qi < p; ++qi) {
        double v[4];
        q1_values(qx[qi], qx[qj], v);
        const double y[2] = {grid.cell_x0(c) + h * qx[qi], grid.cell_y0(c) + h * qx[qj]};
        const double JxW = qw[qi] * qw[qj] * h * h;
        const double fv = f(y);
        for (int i = 0; i < 4; ++i) {
          rhs[cd[i]] += fv * v[i] * JxW;
          for (int j = 0; j < 4; ++j) mass.add(cd[i], cd[j], v[i] * v[j] * JxW);
        }
      }
  }
  double nrm = 0;
  for (double r : rhs) nrm += r * r;
  solve_cg(mass, sol, rhs, 10 * n + 100, 1e-12 * std::sqrt(nrm));
  return sol;
}

// Evaluate a nodal Q1 field at unit-cell point (xi,eta) of cell c.
inline double q1_interp(const Grid& g, const std::vector<double>& u, int c, double xi, double eta) {
  double v[4];
  q1_values(xi, eta, v);
  const auto& cd = g.cell_dofs[c];
  return v[0] * u[cd[0]] + v[1] * u[cd[1]] + v[2] * u[cd[2]] + v[3] * u[cd[3]];
}

// L2 norm over [-1,1]^2 of the nodal Q1 field u with QGauss<2>(p):
// VectorTools::integrate_difference(dof, u, ZeroFunction, cellwise, QGauss(p), L2_norm)
// followed by compute_global_error(L2_norm).
inline double q1_l2_norm(const Grid& g, const std::vector<double>& u, int p) {
  std::vector<double> x, w;
  gauss01(p, x, w);
  double total = 0;
  for (int c = 0; c < g.n_cells(); ++c) {
    double s = 0;
    for (int qy = 0; qy < p; ++qy)
      for (int qx = 0; qx < p; ++qx) {
        double v = q1_interp(g, u, c, x[qx], x[qy]);
        s += v * v * (w[qx] * w[qy] * g.h * g.h);
      }
    double cell = std::sqrt(s);
    total += cell * cell;
  }
  return std::sqrt(total);
}

}  // namespace orc
