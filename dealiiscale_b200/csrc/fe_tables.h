// Host-side tables for FE_Q<2>(1) on a uniform square mesh of [-1,1]^2: Gauss points, the
// per-quadrature-point basis products the moment kernels need, and the reference's DoF
// numbering.
//
// Replaces what the reference gets from deal.II at micro.cpp:141-144 (FEValues + QGauss<2>(12)),
// bio micro.cpp:12-15 (FEFaceValues + QGauss<1>(12)), micro.cpp:41-42 / bio micro.cpp:74
// (mesh generation) and micro.cpp:48 (distribute_dofs).
//
// Device ordering: node (ix,iy) -> r = ix*m + iy and cell (cx,cy) -> c = cx*nc + cy, i.e. the
// y index runs fastest.  With the 32 lanes of a warp on consecutive cy, everything that
// depends on (cx, qx) only is warp-uniform and everything that depends on (cy, qy) is
// unit-stride in memory.
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

#include "kernels.h"

namespace msgpu {

// n-point Gauss-Legendre rule mapped to [0,1]; nodes ascending.
inline void gauss_legendre_unit(int n, std::vector<double>& nodes, std::vector<double>& weights) {
  nodes.resize(n);
  weights.resize(n);
  const long double PI = 3.14159265358979323846264338327950288L;
  for (int k = 0; k < n; ++k) {
    // Chebyshev-like start, then Newton on P_n
    long double x = -cosl(PI * (4 * k + 3) / (4 * n + 2));
    long double dp = 1;
    for (int iter = 0; iter < 64; ++iter) {
      long double pm1 = 1.0L, p = x;  // P_0, P_1
      for (int j = 2; j <= n; ++j) {
        long double pn = ((2 * j - 1) * x * p - (j - 1) * pm1) / j;
        pm1 = p;
        p = pn;
      }
      if (n == 0) p = 1.0L;
      dp = n * (x * p - pm1) / (x * x - 1.0L);
      long double dx = p / dp;
      x -= dx;
      if (fabsl(dx) < 1e-20L) break;
    }
    nodes[k] = static_cast<double>(0.5L * (x + 1.0L));
    weights[k] = static_cast<double>(1.0L / ((1.0L - x * x) * dp * dp));
  }
}

// Per volume quadrature point (qy*q + qx), premultiplied by the unit-cell weight w = w_qx*w_qy:
//   [0..2]  eb^2, eb*e, e^2      (e = eta, eb = 1-eta)   -> d/dxi   * d/dxi   products
//   [3..5]  xb^2, xb*x, x^2      (x = xi,  xb = 1-xi)    -> d/deta  * d/deta  products
//   [6..9]  eb*xb, eb*x, e*xb, e*x                       -> d/dxi   * d/deta  products
//   [10..13] shape values phi_0..phi_3 (deal.II vertex order (0,0),(1,0),(0,1),(1,1))
inline std::vector<double> build_qtab(int q, const std::vector<double>& pt, const std::vector<double>& wt) {
  std::vector<double> t(static_cast<size_t>(q) * q * QTAB_STRIDE);
  for (int qy = 0; qy < q; ++qy)
    for (int qx = 0; qx < q; ++qx) {
      double x = pt[qx], e = pt[qy], xb = 1 - x, eb = 1 - e, w = wt[qx] * wt[qy];
      double* r = &t[(static_cast<size_t>(qy) * q + qx) * QTAB_STRIDE];
      r[0] = w * (eb * eb); r[1] = w * (eb * e); r[2] = w * (e * e);
      r[3] = w * (xb * xb); r[4] = w * (xb * x); r[5] = w * (x * x);
      r[6] = w * (eb * xb); r[7] = w * (eb * x); r[8] = w * (e * xb); r[9] = w * (e * x);
      r[10] = w * (xb * eb); r[11] = w * (x * eb); r[12] = w * (xb * e); r[13] = w * (x * e);
    }
  return t;
}

// 1-D factors of the table above: every entry of build_qtab is (a function of xi) * (a function of eta), so the
// moments of a cell are summed row by row - the xi factor at every point, the eta factor once per row:
//   row 0, 1   w (1 - s), w s                      (shape values; mixed-derivative products)
//   row 2..4   w (1 - s)^2, w (1 - s) s, w s^2     (products of two derivatives in the other direction)
//   row 5      w
// entry [r * q + i] for the i-th Gauss point s = pt[i], w = wt[i].
inline std::vector<double> build_q1d(int q, const std::vector<double>& pt, const std::vector<double>& wt) {
  std::vector<double> t(static_cast<size_t>(Q1D_ROWS) * q);
  for (int i = 0; i < q; ++i) {
    const double s = pt[i], sb = 1 - pt[i], w = wt[i];
    t[0 * q + i] = w * sb;
    t[1 * q + i] = w * s;
    t[2 * q + i] = w * (sb * sb);
    t[3 * q + i] = w * (sb * s);
    t[4 * q + i] = w * (s * s);
    t[5 * q + i] = w;
  }
  return t;
}

// Reference DoF numbering of FE_Q(1): DoFs are numbered in the order the active cells first
// touch their vertices, local vertex order (0,0),(1,0),(0,1),(1,1).  Active cells of
// hyper_cube+refine_global come in Morton order (child index = cx_bit + 2*cy_bit per level),
// those of subdivided_hyper_cube row by row with x fastest.
// Returns ref_to_internal[ref dof] = ix*m + iy.
inline std::vector<int> reference_numbering(int mesh_kind, int nc) {
  const int m = nc + 1;
  std::vector<int> internal_to_ref(static_cast<size_t>(m) * m, -1);
  int levels = 0;
  if (mesh_kind == 0) {
    while ((1 << levels) < nc) ++levels;
    if ((1 << levels) != nc) throw std::runtime_error("refine_global meshes have 2^r cells per axis");
  }
  int next = 0;
  for (int a = 0; a < nc * nc; ++a) {
    int cx, cy;
    if (mesh_kind == 0) {
      cx = cy = 0;
      for (int bit = 0; bit < levels; ++bit) {
        cx |= ((a >> (2 * bit)) & 1) << bit;
        cy |= ((a >> (2 * bit + 1)) & 1) << bit;
      }
    } else {
      cx = a % nc;
      cy = a / nc;
    }
    for (int v = 0; v < 4; ++v) {
      int r = (cx + (v & 1)) * m + (cy + (v >> 1));
      if (internal_to_ref[r] < 0) internal_to_ref[r] = next++;
    }
  }
  std::vector<int> ref_to_internal(static_cast<size_t>(m) * m);
  for (int r = 0; r < m * m; ++r) ref_to_internal[internal_to_ref[r]] = r;
  return ref_to_internal;
}

// parabolic: shared stencils built on the host (n entries, closed-form Q1 local matrices on a
// square cell: what MatrixCreator::create_mass_matrix / create_laplace_matrix with QGauss(2)
// integrate exactly, reference src/elliptic-parabolic/rho_solver.cpp:64-65)
inline void add_local_stencil(std::vector<double>& dg, const MeshDims& md, int cx, int cy, const double loc[4][4], double f) {
  for (int a = 0; a < 4; ++a)
    for (int b = a; b < 4; ++b) {
      const int ra = (cx + (a & 1)) * md.m + (cy + (a >> 1)), rb = (cx + (b & 1)) * md.m + (cy + (b >> 1));
      const int lo = std::min(ra, rb), off = std::abs(rb - ra);
      const int d = off == 0 ? 0 : off == 1 ? 1 : off == md.m - 1 ? 2 : off == md.m ? 3 : 4;
      dg[static_cast<size_t>(d) * md.ld + lo] += f * loc[a][b];
    }
}

inline void build_stencils(const MeshDims& md, double dt, double D, double kappa, double R, std::vector<double>& mass,
                           std::vector<double>& sys, std::vector<double>& lumped) {
  const double h = md.h;
  double M[4][4], L[4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      const int sx = (a & 1) == (b & 1), sy = (a >> 1) == (b >> 1);
      M[a][b] = h * h / 36.0 * (sx ? 2.0 : 1.0) * (sy ? 2.0 : 1.0);
      L[a][b] = (a == b) ? 4.0 / 6.0 : (sx || sy) ? -1.0 / 6.0 : -2.0 / 6.0;
    }
  mass.assign(static_cast<size_t>(NDIAG) * md.ld, 0.0);
  sys.assign(static_cast<size_t>(NDIAG) * md.ld, 0.0);
  lumped.assign(md.n, 0.0);
  for (int cx = 0; cx < md.nc; ++cx)
    for (int cy = 0; cy < md.nc; ++cy) {
      add_local_stencil(mass, md, cx, cy, M, 1.0);
      add_local_stencil(sys, md, cx, cy, M, 1.0);
      add_local_stencil(sys, md, cx, cy, L, dt * D);
      for (int a = 0; a < 4; ++a) lumped[(cx + (a & 1)) * md.m + (cy + (a >> 1))] += h * h / 4.0;
    }
  // Robin mass kappa*dt*R * int phi_i phi_j on the faces |y0| = 1 (rho_solver.cpp:151-178)
  const double c = kappa * dt * R * h / 6.0;
  for (int side = 0; side < 2; ++side) {
    const int ix = side == 0 ? 0 : md.m - 1;
    for (int f = 0; f < md.nc; ++f) {
      const int r = ix * md.m + f;
      sys[r] += 2.0 * c;
      sys[r + 1] += 2.0 * c;
      sys[static_cast<size_t>(1) * md.ld + r] += c;
    }
  }
}


}  // namespace msgpu
