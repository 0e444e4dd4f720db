// Per-thread bodies of the error / residual / point-load kernels (K8 and the parabolic set-up).
// __host__ __device__ for the same reason as assemble_bodies.cuh: tests/native drives them on the
// CPU; the product library launches them on the device only.
//
// Replaces VectorTools::integrate_difference as the reference calls it per micro grid
// (src/elliptic-elliptic/micro.cpp:193-202, src/bio-multiscale/micro.cpp:326-335 and :373-376,
// src/elliptic-parabolic/rho_solver.cpp:315-323) and the right-hand side of VectorTools::project
// (src/elliptic-parabolic/rho_solver.cpp:80-83).
#pragma once
#include "assemble_bodies.cuh"

namespace msgpu {

// One thread per (system, cell).  out[(kl*2 + 0)*ncells + c] = sum_q (v_h - s)^2 JxW,
// out[(kl*2 + 1)*ncells + c] = sum_q |grad v_h - grad s|^2 JxW, with s the exact solution program
// (MODE_POINTS) and grad s by central differences of step fd (AutoDerivativeFunction, h = 1e-8,
// src/tools/multiscale_function_parser.h:44-46) in the REFERENCE-square coordinates.
MSGPU_HOSTDEV void body_cell_errors(const VmProgram& prog, const MeshDims& md, const double* __restrict__ slots,
                                    int n_slots, int qe, const double* __restrict__ pt, const double* __restrict__ wt,
                                    const double* __restrict__ x, int N, double fd, double* __restrict__ out,
                                    long long gtid, double* Rt, int TB) {
  if (gtid >= static_cast<long long>(N) * md.ncells) return;
  const int k = static_cast<int>(gtid / md.ncells), c = static_cast<int>(gtid % md.ncells);
  const int cx = c / md.nc, cy = c % md.nc;
  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.U = slots + static_cast<size_t>(k) * n_slots;
  io.L0 = io.L1 = nullptr;
  io.OUT = nullptr;
  io.l0_stride = io.l1_stride = io.out_stride = 0;
  const double* xk = x + static_cast<size_t>(k) * md.n;
  // nodal values, deal.II vertex order (0,0),(1,0),(0,1),(1,1)
  const double v0 = xk[cx * md.m + cy], v1 = xk[(cx + 1) * md.m + cy];
  const double v2 = xk[cx * md.m + cy + 1], v3 = xk[(cx + 1) * md.m + cy + 1];
  const double ox = -1.0 + cx * md.h, oy = -1.0 + cy * md.h;
  auto eval = [&](double y0, double y1) -> double {
    io.R[0] = y0;
    io.R[TB] = y1;
    vm_run(prog, 0, io);
    return io.R[prog.out0 * TB];
  };
  double l2 = 0.0, semi = 0.0;
  for (int qy = 0; qy < qe; ++qy)
    for (int qx = 0; qx < qe; ++qx) {
      const double xi = pt[qx], eta = pt[qy];
      const double y0 = ox + md.h * xi, y1 = oy + md.h * eta;
      const double jxw = wt[qx] * wt[qy] * md.h * md.h;
      const double vh = v0 * ((1 - xi) * (1 - eta)) + v1 * (xi * (1 - eta)) + v2 * ((1 - xi) * eta) + v3 * (xi * eta);
      const double gx = (-(1 - eta) * v0 + (1 - eta) * v1 - eta * v2 + eta * v3) / md.h;
      const double gy = (-(1 - xi) * v0 - xi * v1 + (1 - xi) * v2 + xi * v3) / md.h;
      const double s = eval(y0, y1);
      const double sx = (eval(y0 + fd, y1) - eval(y0 - fd, y1)) / (2 * fd);
      const double sy = (eval(y0, y1 + fd) - eval(y0, y1 - fd)) / (2 * fd);
      l2 += (vh - s) * (vh - s) * jxw;
      semi += ((gx - sx) * (gx - sx) + (gy - sy) * (gy - sy)) * jxw;
    }
  out[(static_cast<size_t>(k) * 2 + 0) * md.ncells + c] = l2;
  out[(static_cast<size_t>(k) * 2 + 1) * md.ncells + c] = semi;
}

// || x - xold ||^2 over one cell with the exact Q1 mass matrix (what QGauss(12) integrates exactly).
MSGPU_HOSTDEV void body_cell_residual(const MeshDims& md, const double* __restrict__ x,
                                      const double* __restrict__ xold, int N, double* __restrict__ out,
                                      long long gtid) {
  if (gtid >= static_cast<long long>(N) * md.ncells) return;
  const int k = static_cast<int>(gtid / md.ncells), c = static_cast<int>(gtid % md.ncells);
  const int cx = c / md.nc, cy = c % md.nc;
  const size_t o = static_cast<size_t>(k) * md.n;
  const int r0 = cx * md.m + cy, r1 = (cx + 1) * md.m + cy;
  const double e0 = x[o + r0] - xold[o + r0], e1 = x[o + r1] - xold[o + r1];
  const double e2 = x[o + r0 + 1] - xold[o + r0 + 1], e3 = x[o + r1 + 1] - xold[o + r1 + 1];
  const double q = 4.0 * (e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3) + 4.0 * (e0 * e1 + e0 * e2 + e1 * e3 + e2 * e3) +
                   2.0 * (e0 * e3 + e1 * e2);
  out[(static_cast<size_t>(k) * 2 + 0) * md.ncells + c] = q * (md.h * md.h / 36.0);
  out[(static_cast<size_t>(k) * 2 + 1) * md.ncells + c] = 0.0;
}

// Load vector of a MODE_POINTS function with a qe-point rule: cell[(k*CELL_STRIDE + 10 + v)*ncells + c]
// = sum_q f(y_q) phi_v JxW; the other entries of the cell record are left alone (a gather without matrix and
// weight outputs, which is the only consumer, does not read them).
MSGPU_HOSTDEV void body_cell_point_load(const VmProgram& prog, const MeshDims& md, const double* __restrict__ slots,
                                        int n_slots, int qe, const double* __restrict__ pt,
                                        const double* __restrict__ wt, int k0, int nk, double* __restrict__ cell,
                                        long long gtid, double* Rt, int TB) {
  if (gtid >= static_cast<long long>(nk) * md.ncells) return;
  const int kl = static_cast<int>(gtid / md.ncells), c = static_cast<int>(gtid % md.ncells);
  const int cx = c / md.nc, cy = c % md.nc;
  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.U = slots + static_cast<size_t>(k0 + kl) * n_slots;
  io.L0 = io.L1 = nullptr;
  io.OUT = nullptr;
  io.l0_stride = io.l1_stride = io.out_stride = 0;
  const double ox = -1.0 + cx * md.h, oy = -1.0 + cy * md.h;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int qy = 0; qy < qe; ++qy)
    for (int qx = 0; qx < qe; ++qx) {
      const double xi = pt[qx], eta = pt[qy];
      io.R[0] = ox + md.h * xi;
      io.R[TB] = oy + md.h * eta;
      vm_run(prog, 0, io);
      const double f = io.R[prog.out0 * TB] * (wt[qx] * wt[qy] * md.h * md.h);
      a0 += f * ((1 - xi) * (1 - eta));
      a1 += f * (xi * (1 - eta));
      a2 += f * ((1 - xi) * eta);
      a3 += f * (xi * eta);
    }
  double* o = cell + static_cast<size_t>(kl) * CELL_STRIDE * md.ncells + c;
  o[static_cast<size_t>(10) * md.ncells] = a0;
  o[static_cast<size_t>(11) * md.ncells] = a1;
  o[static_cast<size_t>(12) * md.ncells] = a2;
  o[static_cast<size_t>(13) * md.ncells] = a3;
}

}  // namespace msgpu
