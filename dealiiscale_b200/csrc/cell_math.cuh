// Per-quadrature-point and per-cell arithmetic of the batched micro assembly (kernel K1).
//
// Replaces the innermost loops of MicroSolver::integrate_cell (reference:
// src/elliptic-elliptic/micro.cpp:117-135) and of local_assemble_system
// (src/bio-multiscale/micro.cpp:201-221) together with compute_pullback_objects (:85-90).
//
// For FE_Q(1) on a square cell of side h the reference-cell gradients are
//   d/dxi  phi = (-eb, +eb, -e, +e),   d/deta phi = (-xb, -x, +xb, +x)      (e=eta, x=xi, b = 1-.)
// and grad phi = (1/h) * that, so the h^2 of JxW cancels in the stiffness entries:
//   A[a][b] = sum_q w_q det(F) ( K00 dxi_a dxi_b + K01 (dxi_a deta_b + deta_a dxi_b) + K11 deta_a deta_b )
// Every A[a][b] is a signed sum of TEN moments (3 for K00, 3 for K11, 4 for K01), so a
// quadrature point costs 10 FMAs instead of the 16x(2x2 tensor contraction) of the reference.
#pragma once
#include "expr_vm.cuh"

namespace msgpu {

struct PullBack {
  double det, k00, k01, k11;
};

// F row-major.  invert(F) as deal.II does it (adjugate times 1/det), K = F^-1 F^-T
// (micro.cpp:87-89).
MSGPU_HOSTDEV PullBack pullback(double f00, double f01, double f10, double f11) {
  PullBack p;
  p.det = f00 * f11 - f10 * f01;
  const double inv_det = 1.0 / p.det;
  const double i00 = f11 * inv_det, i01 = -f01 * inv_det, i10 = -f10 * inv_det, i11 = f00 * inv_det;
  p.k00 = i00 * i00 + i01 * i01;
  p.k01 = i00 * i10 + i01 * i11;
  p.k11 = i10 * i10 + i11 * i11;
  return p;
}

// Symmetric 4x4 local matrix stored as (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3).
// mom: [0..2] K00 moments (eb^2, eb e, e^2), [3..5] K11 moments (xb^2, xb x, x^2),
//      [6..9] K01 moments (eb xb, eb x, e xb, e x).
MSGPU_HOSTDEV void moments_to_local(const double* mom, double scale, double* a) {
  const double A0 = mom[0], A1 = mom[1], A2 = mom[2];
  const double C0 = mom[3], C1 = mom[4], C2 = mom[5];
  const double B00 = mom[6], B01 = mom[7], B10 = mom[8], B11 = mom[9];
  a[0] = scale * (A0 + C0 + 2.0 * B00);
  a[1] = scale * (-A0 + C1 + B01 - B00);
  a[2] = scale * (A1 - C0 - B00 + B10);
  a[3] = scale * (-A1 - C1 - B01 - B10);
  a[4] = scale * (A0 + C2 - 2.0 * B01);
  a[5] = scale * (-A1 - C1 + B00 + B11);
  a[6] = scale * (A1 - C2 + B01 - B11);
  a[7] = scale * (A2 + C0 - 2.0 * B10);
  a[8] = scale * (-A2 + C1 - B11 + B10);
  a[9] = scale * (A2 + C2 + 2.0 * B11);
}

// index of entry (a,b), a<=b, in the packed symmetric local matrix
MSGPU_HOSTDEV int sym4(int a, int b) {
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * 4 - (lo * (lo - 1)) / 2 + (hi - lo);
}

// Length of F * Rot * n for the outward normal of a side of the reference square, with
// Rot = [[0,-1],[1,0]] (src/tools/pde_data.cpp:205-206, pde_data.h:68-71):
// sides x = -1 / +1 -> |(F01,F11)|, sides y = -1 / +1 -> |(F00,F10)|.
MSGPU_HOSTDEV double boundary_stretch(int side_is_x_const, double f00, double f01, double f10, double f11) {
  return side_is_x_const ? sqrt(f01 * f01 + f11 * f11) : sqrt(f00 * f00 + f10 * f10);
}

}  // namespace msgpu
