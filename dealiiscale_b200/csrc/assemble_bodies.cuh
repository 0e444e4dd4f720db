// Per-thread bodies of the assembly kernels (K0, K1, K2).  Each body is the complete work of one
// CUDA thread; kernels_assemble.cu wraps them in __global__ kernels.  They are __host__ __device__
// so that tests/native can drive the very same code sequentially on the CPU and compare it with
// the oracle when no GPU is available — the product library only launches them on the device.
//
// Replaces, for all N micro systems at once,
//   MicroSolver::compute_pullback_objects + integrate_cell + assemble_system
//       (reference: src/elliptic-elliptic/micro.cpp:75-171)
//   MicroSolver::local_assemble_system + copy_local_to_global
//       (reference: src/bio-multiscale/micro.cpp:183-288)
#pragma once
#include "cell_math.cuh"
#include "expr_vm.cuh"
#include "kernels.h"

#ifdef __CUDA_ARCH__
#define MSGPU_LDG(p) __ldg(p)
#else
#define MSGPU_LDG(p) (*(p))
#endif

namespace msgpu {

// ---------------------------------------------------------------- K0: per-system slots
MSGPU_HOSTDEV void body_slots(const VmProgram& prog, const double* __restrict__ xy, int N, double t,
                              double* __restrict__ slots, int n_slots, long long gtid, double* Rt, int TB) {
  const int k = static_cast<int>(gtid);
  if (k >= N) return;
  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.R[0 * TB] = xy[2 * k];
  io.R[1 * TB] = xy[2 * k + 1];
  io.R[2 * TB] = t;
  io.OUT = slots + static_cast<size_t>(k) * n_slots;
  io.out_stride = 1;
  io.U = io.OUT;  // slots written earlier in the program are read back through U
  io.L0 = io.L1 = nullptr;
  io.l0_stride = io.l1_stride = 0;
  vm_run(prog, 0, io);
}

// ---------------------------------------------------------------- K0: line tables
MSGPU_HOSTDEV void body_lines(const VmProgram& prog0, const VmProgram& prog1, const double* __restrict__ coord0,
                              const double* __restrict__ coord1, int P, int N, const double* __restrict__ slots,
                              int n_slots, double* __restrict__ T0, int n0, double* __restrict__ T1, int n1,
                              long long gtid, double* Rt, int TB) {
  const long long idx = gtid;
  if (idx >= static_cast<long long>(N) * P) return;
  const int k = static_cast<int>(idx / P), p = static_cast<int>(idx % P);
  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.U = slots + static_cast<size_t>(k) * n_slots;
  io.out_stride = P;
  io.l0_stride = io.l1_stride = P;
  if (n0 > 0) {
    io.R[0] = coord0[p];
    io.OUT = T0 + static_cast<size_t>(k) * n0 * P + p;
    io.L0 = io.OUT;
    io.L1 = nullptr;
    vm_run(prog0, 0, io);
  }
  if (n1 > 0) {
    io.R[0] = coord1[p];
    io.OUT = T1 + static_cast<size_t>(k) * n1 * P + p;
    io.L1 = io.OUT;
    io.L0 = nullptr;
    vm_run(prog1, 0, io);
  }
}

// ---------------------------------------------------------------- K1: cell moments
// One thread per (system, cell); q*q quadrature points in registers, no reduction across threads.
MSGPU_HOSTDEV void body_cell_moments(const VmProgram& prog, const MeshDims& md, const Tables& tb,
                                     const double* __restrict__ qtab, int k0, int nk, double stiffness_scale,
                                     int has_map, double* __restrict__ cell, int* __restrict__ error_flag,
                                     long long gtid, double* Rt, int TB) {
  const long long t = gtid;
  if (t >= static_cast<long long>(nk) * md.ncells) return;
  const int kl = static_cast<int>(t / md.ncells), c = static_cast<int>(t % md.ncells);
  const int cx = c / md.nc, cy = c % md.nc;
  const int k = k0 + kl;

  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.U = tb.slots + static_cast<size_t>(k) * tb.n_slots;
  io.l0_stride = io.l1_stride = md.P;
  io.OUT = nullptr;
  io.out_stride = 0;
  const double* T0k = tb.T0 + static_cast<size_t>(k) * tb.n0 * md.P + cx * md.q;
  const double* T1k = tb.T1 + static_cast<size_t>(k) * tb.n1 * md.P + cy;
  const double* out = io.R + prog.out0 * TB;

  double acc[CELL_STRIDE];
#pragma unroll
  for (int i = 0; i < CELL_STRIDE; ++i) acc[i] = 0.0;
  bool bad = false;

  for (int qy = 0; qy < md.q; ++qy) {
    io.L1 = T1k + qy * md.nc;
    for (int qx = 0; qx < md.q; ++qx) {
      io.L0 = T0k + qx;
      vm_run(prog, 0, io);
      const double* tab = qtab + (qy * md.q + qx) * QTAB_STRIDE;
      double det, m00, m01, m11, g;
      if (has_map) {
        const PullBack pb = pullback(out[0 * TB], out[1 * TB], out[2 * TB], out[3 * TB]);
        g = out[4 * TB];
        det = pb.det;
        m00 = pb.k00 * det;
        m01 = pb.k01 * det;
        m11 = pb.k11 * det;
        bad |= !(det > 0.0);
      } else {
        g = out[0];
        det = 1.0;
        m00 = 1.0;
        m01 = 0.0;
        m11 = 1.0;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) acc[i] += m00 * MSGPU_LDG(tab + i);
#pragma unroll
      for (int i = 3; i < 6; ++i) acc[i] += m11 * MSGPU_LDG(tab + i);
#pragma unroll
      for (int i = 6; i < 10; ++i) acc[i] += m01 * MSGPU_LDG(tab + i);
      const double gd = g * det;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double phi = MSGPU_LDG(tab + 10 + i);
        acc[10 + i] += gd * phi;
        acc[14 + i] += det * phi;
      }
    }
  }
  double a[10];
  moments_to_local(acc, stiffness_scale, a);
  const double h2 = md.h * md.h;
  double* o = cell + static_cast<size_t>(kl) * CELL_STRIDE * md.ncells + c;
#pragma unroll
  for (int e = 0; e < 10; ++e) o[static_cast<size_t>(e) * md.ncells] = a[e];
#pragma unroll
  for (int e = 10; e < CELL_STRIDE; ++e) o[static_cast<size_t>(e) * md.ncells] = acc[e] * h2;
  if (bad) *error_flag = 1;  // benign race: every writer stores the same value
}


// ---------------------------------------------------------------- K1: cell moments, vectorised
// Same result as body_cell_moments, organised for speed:
//  * the VM runs once per row of W quadrature points (vm_run_vec) instead of once per point;
//  * MODE_GENERAL : F depends on y            -> program `prog` yields F00 F01 F10 F11 G per point
//    MODE_CONST_F : F constant in the cell    -> `progF` runs once, `prog` yields G per point and
//                                                the ten stiffness moments are (K det) * sum_q(tab)
//    MODE_NO_MAP  : no mapping (F = identity) -> as MODE_CONST_F with K = I, det = 1
//  * Tab is an accessor of the per-point basis products (constant memory on the device).
enum MomentMode : int { MODE_GENERAL = 0, MODE_CONST_F = 1, MODE_NO_MAP = 2 };

struct TabGlobal {
  const double* tab;
  const double* sum;
  const double* q1d;  // build_q1d: Q1D_ROWS rows of nq 1-D factors
  int nq;
  MSGPU_HOSTDEV double operator()(int q, int i) const { return MSGPU_LDG(tab + q * QTAB_STRIDE + i); }
  MSGPU_HOSTDEV double total(int i) const { return MSGPU_LDG(sum + i); }
  MSGPU_HOSTDEV double line(int r, int i) const { return MSGPU_LDG(q1d + r * nq + i); }
};

template <int Q, int W, int MODE, typename Tab>
MSGPU_HOSTDEV void body_cell_moments_vec(const VmProgram& progF, const VmProgram& prog, const MeshDims& md,
                                         const Tables& tb, const Tab& tab, int k0, int nk, double stiffness_scale,
                                         double* __restrict__ cell, int* __restrict__ error_flag, long long gtid,
                                         double* Rt, int TB) {
  static_assert(Q % W == 0, "row of quadrature points must split into whole vectors");
  const long long t = gtid;
  if (t >= static_cast<long long>(nk) * md.ncells) return;
  const int kl = static_cast<int>(t / md.ncells), c = static_cast<int>(t % md.ncells);
  const int cx = c / md.nc, cy = c % md.nc;
  const int k = k0 + kl;

  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.U = tb.slots + static_cast<size_t>(k) * tb.n_slots;
  io.l0_stride = io.l1_stride = md.P;
  io.OUT = nullptr;
  io.out_stride = 0;
  const double* T0k = tb.T0 + static_cast<size_t>(k) * tb.n0 * md.P + cx * Q;
  const double* T1k = tb.T1 + static_cast<size_t>(k) * tb.n1 * md.P + cy;

  double acc[CELL_STRIDE];
#pragma unroll
  for (int i = 0; i < CELL_STRIDE; ++i) acc[i] = 0.0;
  bool bad = false;
  double det_c = 1.0, m00_c = 1.0, m01_c = 0.0, m11_c = 1.0;
  if (MODE == MODE_CONST_F) {
    // F only reads constants and system slots here, so the table pointers are irrelevant
    io.L0 = T0k;
    io.L1 = T1k;
    vm_run(progF, 0, io);
    const double* f = io.R + progF.out0 * TB;
    const PullBack pb = pullback(f[0 * TB], f[1 * TB], f[2 * TB], f[3 * TB]);
    det_c = pb.det;
    m00_c = pb.k00 * det_c;
    m01_c = pb.k01 * det_c;
    m11_c = pb.k11 * det_c;
    bad |= !(det_c > 0.0);
  }
  const double* out = io.R + static_cast<size_t>(prog.out0) * W * TB;  // output j, lane e at out[(j*W + e)*TB]

  // Every table entry is (a factor in xi) * (a factor in eta), fe_tables.h build_q1d: sums over a row of points with
  // the xi factors, the eta factors once per row (constant F: 2 instead of 5 FP64 operations per point, general F:
  // 11 instead of 19 besides the pull-back).  The generated kernel (jit_codegen.cpp) does the same, statement by
  // statement.
  for (int qy = 0; qy < Q; ++qy) {
    io.L1 = T1k + qy * md.nc;
    double sa = 0.0, sc0 = 0.0, sc1 = 0.0, sc2 = 0.0, sb0 = 0.0, sb1 = 0.0, sd0 = 0.0, sd1 = 0.0, sg0 = 0.0, sg1 = 0.0;
#pragma unroll
    for (int ch = 0; ch < Q / W; ++ch) {
      io.L0 = T0k + ch * W;
      vm_run_vec<W>(prog, 0, io);
#pragma unroll
      for (int e = 0; e < W; ++e) {
        const int qx = ch * W + e;
        if (MODE == MODE_GENERAL) {
          const PullBack pb = pullback(out[(0 * W + e) * TB], out[(1 * W + e) * TB], out[(2 * W + e) * TB],
                                       out[(3 * W + e) * TB]);
          const double g = out[(4 * W + e) * TB];
          const double det = pb.det;
          const double m00 = pb.k00 * det, m01 = pb.k01 * det, m11 = pb.k11 * det;
          bad |= !(det > 0.0);
          sa += m00 * tab.line(5, qx);
          sc0 += m11 * tab.line(2, qx);
          sc1 += m11 * tab.line(3, qx);
          sc2 += m11 * tab.line(4, qx);
          sb0 += m01 * tab.line(0, qx);
          sb1 += m01 * tab.line(1, qx);
          sd0 += det * tab.line(0, qx);
          sd1 += det * tab.line(1, qx);
          const double gd = g * det;
          sg0 += gd * tab.line(0, qx);
          sg1 += gd * tab.line(1, qx);
        } else {
          const double g = out[e * TB];
          sg0 += g * tab.line(0, qx);
          sg1 += g * tab.line(1, qx);
        }
      }
    }
    const double b0 = tab.line(0, qy), b1 = tab.line(1, qy);
    if (MODE == MODE_GENERAL) {
      acc[0] += sa * tab.line(2, qy);
      acc[1] += sa * tab.line(3, qy);
      acc[2] += sa * tab.line(4, qy);
      acc[3] += sc0 * tab.line(5, qy);
      acc[4] += sc1 * tab.line(5, qy);
      acc[5] += sc2 * tab.line(5, qy);
      acc[6] += sb0 * b0;
      acc[7] += sb1 * b0;
      acc[8] += sb0 * b1;
      acc[9] += sb1 * b1;
      acc[14] += sd0 * b0;
      acc[15] += sd1 * b0;
      acc[16] += sd0 * b1;
      acc[17] += sd1 * b1;
    }
    acc[10] += sg0 * b0;
    acc[11] += sg1 * b0;
    acc[12] += sg0 * b1;
    acc[13] += sg1 * b1;
  }
  if (MODE != MODE_GENERAL) {
#pragma unroll
    for (int i = 10; i < 14; ++i) acc[i] *= det_c;
  }
  if (MODE != MODE_GENERAL) {
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] = m00_c * tab.total(i);
#pragma unroll
    for (int i = 3; i < 6; ++i) acc[i] = m11_c * tab.total(i);
#pragma unroll
    for (int i = 6; i < 10; ++i) acc[i] = m01_c * tab.total(i);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[14 + i] = det_c * tab.total(10 + i);
  }
  double a[10];
  moments_to_local(acc, stiffness_scale, a);
  const double h2 = md.h * md.h;
  double* o = cell + static_cast<size_t>(kl) * CELL_STRIDE * md.ncells + c;
  if (MODE != MODE_GENERAL) {
    // F is constant per system: the matrix entries and the det-weighted basis integrals are the same in every
    // cell and the gather reads them from cell 0 (GatherArgs::uniform_stiffness) — as the specialised kernel does
    if (c == 0) {
#pragma unroll
      for (int e = 0; e < 10; ++e) o[static_cast<size_t>(e) * md.ncells] = a[e];
#pragma unroll
      for (int e = 14; e < CELL_STRIDE; ++e) o[static_cast<size_t>(e) * md.ncells] = acc[e] * h2;
    }
#pragma unroll
    for (int e = 10; e < 14; ++e) o[static_cast<size_t>(e) * md.ncells] = acc[e] * h2;
  } else {
#pragma unroll
    for (int e = 0; e < 10; ++e) o[static_cast<size_t>(e) * md.ncells] = a[e];
#pragma unroll
    for (int e = 10; e < CELL_STRIDE; ++e) o[static_cast<size_t>(e) * md.ncells] = acc[e] * h2;
  }
  if (bad) *error_flag = 1;  // benign race: every writer stores the same value
}

// ---------------------------------------------------------------- K1: face moments
// One thread per (system, boundary face of one side).
MSGPU_HOSTDEV void body_face_moments(const VmProgram& prog, int side, const MeshDims& md, const Tables& tb,
                                     const double* __restrict__ q1d_pt, const double* __restrict__ q1d_wt, int k0,
                                     int nk, int has_map, double* __restrict__ face, long long gtid, double* Rt,
                                     int TB) {
  const long long t = gtid;
  if (t >= static_cast<long long>(nk) * md.nc) return;
  const int kl = static_cast<int>(t / md.nc), f = static_cast<int>(t % md.nc);
  const int k = k0 + kl;
  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.U = tb.slots + static_cast<size_t>(k) * tb.n_slots;
  io.l0_stride = io.l1_stride = md.P;
  io.OUT = nullptr;
  io.out_stride = 0;
  const double* T0k = tb.T0 + static_cast<size_t>(k) * tb.n0 * md.P;
  const double* T1k = tb.T1 + static_cast<size_t>(k) * tb.n1 * md.P;
  const int vtx = md.nc * md.q;  // first vertex entry of the 1-D point list
  const double* out = io.R + prog.out0 * TB;
  double acc[FACE_STRIDE];
#pragma unroll
  for (int i = 0; i < FACE_STRIDE; ++i) acc[i] = 0.0;
  for (int qf = 0; qf < md.q; ++qf) {
    if (side < 2) {  // x = -1 or x = +1, running along y
      io.L0 = T0k + vtx + (side == 0 ? 0 : md.m - 1);
      io.L1 = T1k + qf * md.nc + f;
    } else {         // y = -1 or y = +1, running along x
      io.L0 = T0k + f * md.q + qf;
      io.L1 = T1k + vtx + (side == 2 ? 0 : md.m - 1);
    }
    // a program without line-table operands (constant boundary data on a map that does not depend on y: every
    // bio set without a manufactured solution) gives the same outputs at every point: one run per face
    if (qf == 0 || prog.reads_lines) vm_run(prog, 0, io);
    double jb = 1.0, bc;
    if (has_map) {
      jb = boundary_stretch(side < 2, out[0 * TB], out[1 * TB], out[2 * TB], out[3 * TB]);
      bc = out[4 * TB];
    } else {
      bc = out[0];
    }
    const double s = q1d_pt[qf], sb = 1.0 - s;
    const double w = q1d_wt[qf] * md.h * jb;
    acc[0] += w * sb * sb;
    acc[1] += w * sb * s;
    acc[2] += w * s * s;
    acc[3] += w * bc * sb;
    acc[4] += w * bc * s;
    acc[5] += w * sb;
    acc[6] += w * s;
  }
  double* o = face + (static_cast<size_t>(kl) * 4 + side) * FACE_STRIDE * md.nc + f;
#pragma unroll
  for (int e = 0; e < FACE_STRIDE; ++e) o[static_cast<size_t>(e) * md.nc] = acc[e];
}

// ---------------------------------------------------------------- K1: gather
// One thread per (system, node): sums the contributions of the <= 4 cells around the node
// in the order (cx-1,cy-1), (cx,cy-1), (cx-1,cy), (cx,cy) — no atomics, deterministic.
// body_gather_node: system kl of the chunk, node r.  UNIFORM (GatherArgs::uniform_stiffness): the ten matrix entries
// and the four weights of a cell are those of cell 0.  Indices inside a system are 32-bit (18 * ncells entries), the
// per-system offsets 64-bit; the sums keep the order above whatever the template argument.  (Round 2: the kernel
// was bound by instruction issue - 460 instructions per node, most of them 64-bit index arithmetic and a 64-bit
// division of the global thread index - not by the 1.5 GB it moves.)
template <bool UNIFORM>
MSGPU_HOSTDEV void body_gather_node(const GatherArgs& a, int kl, int r) {
  const MeshDims& md = a.md;
  const int nc = md.nc, nce = md.ncells;
  const int ix = r / md.m, iy = r - ix * md.m;
  const int k = a.k0 + kl;
  const double* const cellk = a.cell + static_cast<size_t>(kl) * CELL_STRIDE * nce;
  const int c11 = ix * nc + iy;  // cell (ix, iy); (ix-1, iy): c11 - nc, (ix, iy-1): c11 - 1, (ix-1, iy-1): c11 - nc - 1
  // entry e of the cell at offset `off` from c11 (matrix entries and weights of a uniform system: cell 0)
  auto C = [&](int off, int e) -> double {
    const int c = (UNIFORM && (e < 10 || e >= 14)) ? 0 : c11 + off;
    return MSGPU_LDG(cellk + (e * nce + c));
  };
  // the parabolic driver and the L2 projections only want the load vector: the matrix is shared and built
  // on the host, so the ten matrix entries and the four weights of a cell are not even read
  const bool want_matrix = a.diag != nullptr, want_jw = a.jw != nullptr;
  const bool right_ok = ix < nc, up_ok = iy < nc, down_ok = iy > 0, left_ok = ix > 0;
  double d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0, load = 0, jw = 0;
  // cells (ix-1, iy-1), (ix, iy-1), (ix-1, iy), (ix, iy): this node is their local vertex 3, 2, 1, 0
  if (left_ok && down_ok) {
    if (want_matrix) d0 += C(-nc - 1, sym4(3, 3));
    load += C(-nc - 1, 10 + 3);
    if (want_jw) jw += C(-nc - 1, 14 + 3);
  }
  if (right_ok && down_ok) {
    if (want_matrix) d0 += C(-1, sym4(2, 2));
    load += C(-1, 10 + 2);
    if (want_jw) jw += C(-1, 14 + 2);
  }
  if (left_ok && up_ok) {
    if (want_matrix) d0 += C(-nc, sym4(1, 1));
    load += C(-nc, 10 + 1);
    if (want_jw) jw += C(-nc, 14 + 1);
  }
  if (right_ok && up_ok) {
    if (want_matrix) d0 += C(0, sym4(0, 0));
    load += C(0, 10 + 0);
    if (want_jw) jw += C(0, 14 + 0);
  }
  if (want_matrix && up_ok) {  // edge to (ix, iy+1)
    if (left_ok) d1 += C(-nc, 6);   // local (1,3)
    if (right_ok) d1 += C(0, 2);    // local (0,2)
  }
  if (want_matrix && right_ok) {  // edge to (ix+1, iy)
    if (down_ok) d3 += C(-1, 8);   // local (2,3)
    if (up_ok) d3 += C(0, 1);      // local (0,1)
    if (up_ok) d4 += C(0, 3);      // diagonal to (ix+1, iy+1): local (0,3)
    if (down_ok) d2 += C(-1, 5);   // diagonal to (ix+1, iy-1): local (1,2)
  }
  if (a.face != nullptr && (ix == 0 || ix == md.m - 1 || iy == 0 || iy == md.m - 1)) {
    const double* facek = a.face + static_cast<size_t>(kl) * 4 * FACE_STRIDE * nc;
    auto F = [&](int side, int e, int f) -> double { return MSGPU_LDG(facek + ((side * FACE_STRIDE + e) * nc + f)); };
    double* edge = a.edge + static_cast<size_t>(k) * 4 * md.m;
    if (ix == 0 || ix == md.m - 1) {
      const int side = ix == 0 ? 0 : 1;
      const double kap = ix == 0 ? a.robin_left : a.robin_right;
      double mass = 0, bcl = 0, sw = 0;
      if (down_ok) { mass += F(side, 2, iy - 1); bcl += F(side, 4, iy - 1); sw += F(side, 6, iy - 1); }
      if (up_ok)   { mass += F(side, 0, iy);     bcl += F(side, 3, iy);     sw += F(side, 5, iy);
                     d1 += kap * F(side, 1, iy); }
      d0 += kap * mass;
      load += bcl;
      edge[side * md.m + iy] = sw;
      edge[(2 + side) * md.m + iy] = bcl;
    }
    if (iy == 0 || iy == md.m - 1) {
      const int side = iy == 0 ? 2 : 3;
      if (left_ok) load += F(side, 4, ix - 1);
      if (right_ok) load += F(side, 3, ix);
    }
  }
  if (want_matrix) {  // parabolic: all systems share one matrix, built once on the host
    double* dg = a.diag + (static_cast<size_t>(k) * NDIAG * md.ld + r);
    dg[0] = d0;
    dg[md.ld] = d1;
    dg[2 * md.ld] = d2;
    dg[3 * md.ld] = d3;
    dg[4 * md.ld] = d4;
  }
  a.b0[static_cast<size_t>(k) * md.n + r] = load;
  if (want_jw) a.jw[static_cast<size_t>(k) * md.n + r] = jw;
}

// global thread index -> (system, node): the sequential CPU harness (tests/native) walks through these
MSGPU_HOSTDEV void body_gather(const GatherArgs& a, long long gtid) {
  const MeshDims& md = a.md;
  if (gtid >= static_cast<long long>(a.nk) * md.n) return;
  const int kl = static_cast<int>(gtid / md.n), r = static_cast<int>(gtid % md.n);
  if (a.uniform_stiffness != 0)
    body_gather_node<true>(a, kl, r);
  else
    body_gather_node<false>(a, kl, r);
}

// ---------------------------------------------------------------- K2: Dirichlet values
MSGPU_HOSTDEV void body_boundary_values(const VmProgram& prog, const MeshDims& md, const Tables& tb, int N,
                                        double* __restrict__ bv, long long gtid, double* Rt, int TB) {
  const int nb = 4 * md.nc;  // boundary nodes of the square
  const long long t = gtid;
  if (t >= static_cast<long long>(N) * nb) return;
  const int k = static_cast<int>(t / nb), j = static_cast<int>(t % nb);
  int ix, iy;  // walk the four sides
  if (j < md.nc) { ix = 0; iy = j; }
  else if (j < 2 * md.nc) { ix = j - md.nc; iy = md.nc; }
  else if (j < 3 * md.nc) { ix = md.nc; iy = md.nc - (j - 2 * md.nc); }
  else { ix = md.nc - (j - 3 * md.nc); iy = 0; }
  VmIo io;
  io.R = Rt;
  io.r_stride = TB;
  io.U = tb.slots + static_cast<size_t>(k) * tb.n_slots;
  io.l0_stride = io.l1_stride = md.P;
  io.OUT = nullptr;
  io.out_stride = 0;
  const int vtx = md.nc * md.q;
  io.L0 = tb.T0 + static_cast<size_t>(k) * tb.n0 * md.P + vtx + ix;
  io.L1 = tb.T1 + static_cast<size_t>(k) * tb.n1 * md.P + vtx + iy;
  vm_run(prog, 0, io);
  bv[static_cast<size_t>(k) * md.n + ix * md.m + iy] = io.R[prog.out0 * TB];
}

MSGPU_HOSTDEV bool on_boundary(int ix, int iy, int m) {
  return ix == 0 || iy == 0 || ix == m - 1 || iy == m - 1;
}

// ---------------------------------------------------------------- K2: right-hand side + elimination
// MatrixTools::apply_boundary_values (micro.cpp:168): boundary row d keeps its diagonal a_dd,
// b_d = v_d a_dd, x_d = v_d; every row r coupled to d gets b_r -= a_rd / a_dd * (v_d a_dd).
// The couplings themselves are zeroed by k_mask_dirichlet afterwards.
// (system k, node r: the kernels run on a (system, node-block) grid, the CPU harness walks through global indices)
MSGPU_HOSTDEV void body_rhs_node(const RhsArgs& a, int k, int r);
MSGPU_HOSTDEV void body_rhs(const RhsArgs& a, long long gtid) {
  const MeshDims& md = a.md;
  if (gtid >= static_cast<long long>(a.N) * md.n) return;
  body_rhs_node(a, static_cast<int>(gtid / md.n), static_cast<int>(gtid % md.n));
}
MSGPU_HOSTDEV void body_rhs_node(const RhsArgs& a, int k, int r) {
  const MeshDims& md = a.md;
  const int ix = r / md.m, iy = r - ix * md.m;
  const size_t vo = static_cast<size_t>(k) * md.n;
  double val = a.b0[vo + r];
  if (a.problem == PF_ELLIPTIC) {
    val += a.u[k] * a.jw[vo + r];
    const double* dg = a.diag + static_cast<size_t>(k) * NDIAG * md.ld;
    const double* bv = a.bv + vo;
    if (on_boundary(ix, iy, md.m)) {
      val = bv[r] * dg[r];
      a.x[vo + r] = bv[r];
    } else {
      const int off[4] = {1, md.m - 1, md.m, md.m + 1};
      const int ddx[4] = {0, 1, 1, 1}, ddy[4] = {1, -1, 0, 1};
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        // lower neighbour r - off[d], coupling stored at its row
        int jx = ix - ddx[d], jy = iy - ddy[d];
        if (on_boundary(jx, jy, md.m)) {
          const int rb = r - off[d];
          const double ard = dg[static_cast<size_t>(d + 1) * md.ld + rb], add = dg[rb];
          val -= ard / add * (bv[rb] * add);
        }
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        int jx = ix + ddx[d], jy = iy + ddy[d];
        if (on_boundary(jx, jy, md.m)) {
          const int rb = r + off[d];
          const double ard = dg[static_cast<size_t>(d + 1) * md.ld + r], add = dg[rb];
          val -= ard / add * (bv[rb] * add);
        }
      }
      a.x[vo + r] = 0.0;
    }
  } else if (a.problem == PF_BIO) {
    const double* edge = a.edge + static_cast<size_t>(k) * 4 * md.m;
    if (ix == 0) val += a.coef_u * a.u[k] * edge[iy];
    if (ix == md.m - 1) val += a.coef_w * a.w[k] * edge[md.m + iy];
  } else {
    // RhoSolver::assemble_system (rho_solver.cpp:138-147, :180-216) with implicit Euler:
    //   b = M rho_old + dt * (volume load + Robin/Neumann face loads) + dt*kappa*(pi_k + p_F) * int_Robin phi
    // a.b holds M rho_old on entry; the start vector is zero (rho_solver.cpp:136).
    const double* edge = a.edge + static_cast<size_t>(k) * 4 * md.m;
    val = a.b[vo + r] + a.dt * val;
    const double robin = a.dt * a.coef_u * (a.u[k] + a.coef_w);
    if (ix == 0) val += robin * edge[iy];
    if (ix == md.m - 1) val += robin * edge[md.m + iy];
    a.x[vo + r] = 0.0;
  }
  a.b[vo + r] = val;
}

MSGPU_HOSTDEV void body_mask_dirichlet_node(const MeshDims& md, double* __restrict__ diag, int k, int r);
MSGPU_HOSTDEV void body_mask_dirichlet(const MeshDims& md, int N, double* __restrict__ diag, long long gtid) {
  if (gtid >= static_cast<long long>(N) * md.n) return;
  body_mask_dirichlet_node(md, diag, static_cast<int>(gtid / md.n), static_cast<int>(gtid % md.n));
}
MSGPU_HOSTDEV void body_mask_dirichlet_node(const MeshDims& md, double* __restrict__ diag, int k, int r) {
  const int ix = r / md.m, iy = r - ix * md.m;
  double* dg = diag + static_cast<size_t>(k) * NDIAG * md.ld;
  const int ddx[4] = {0, 1, 1, 1}, ddy[4] = {1, -1, 0, 1};
  const bool me = on_boundary(ix, iy, md.m);
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int jx = ix + ddx[d], jy = iy + ddy[d];
    const bool inside = jx >= 0 && jy >= 0 && jx < md.m && jy < md.m;
    if (inside && (me || on_boundary(jx, jy, md.m))) dg[static_cast<size_t>(d + 1) * md.ld + r] = 0.0;
  }
}


}  // namespace msgpu
