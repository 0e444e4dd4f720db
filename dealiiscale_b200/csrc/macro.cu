// The macro system(s) on the device: the step on the other side of the
// coupling allgather, so that a fixed-point cycle needs no host round trip.
//
// Replaces MacroSolver::assemble_system + solve of the reference
//   src/elliptic-elliptic/macro.cpp:57-114   (one field, Dirichlet on the whole boundary)
//   src/bio-multiscale/macro.cpp:100-202     (fields u and w)
//   src/elliptic-parabolic/pi_solver.cpp:76-149 (the solve, and the part of its load that changes with the micro
//                                               solutions: int phi_i rho_h pi_h, k_macro_product)
// The reference re-assembles the macro matrix in every cycle although it never changes, and its
// right-hand side is AFFINE in the micro contribution c (macro.cpp:78-93: the interpolated
// contribution enters the load linearly).  The host therefore hands over, once,
//   A   the matrix after MatrixTools::apply_boundary_values (symmetric 9-point stencil),
//   b0  the load for c = 0 after boundary elimination,
//   C   the 9-point operator with  b(c) = b0 + C c  (rows of Dirichlet DoFs are zero),
// and every cycle is: k_macro_rhs (b = b0 + C c, c read from the device-resident gather buffer of
// msgpu_coupling) -> the batched CG kernels of kernels_cg.cu with one system per field (same
// recurrence, warm start as in the reference) -> k_macro_adopt (the solution becomes u_k / w_k of
// the micro systems of this rank).  The macro mesh is the mesh of the grid locations: macro DoF i
// (reference numbering) is micro system i.
#include <cuda_runtime.h>

#include <cmath>
#include <string>
#include <vector>

#include "../../include/msgpu.h"
#include "ctx.h"
#include "fe_tables.h"
#include "kernels.h"

namespace msgpu {
int comm_block(const msgpu_ctx* ctx);
}
using namespace msgpu;

namespace {

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
      return MSGPU_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

int fail(msgpu_ctx* ctx, int code, const std::string& msg) {
  ctx->err = msg;
  return code;
}

constexpr int NOFF = 9;  // general 9-point operator: offsets 0, +1, -1, +(m-1), -(m-1), +m, -m, +(m+1), -(m+1)

__host__ __device__ inline int stencil_offset(int o, int m) {
  const int base = o == 0 ? 0 : (o + 1) / 2 == 1 ? 1 : (o + 1) / 2 == 2 ? m - 1 : (o + 1) / 2 == 3 ? m : m + 1;
  return (o & 1) || o == 0 ? base : -base;
}

// b[f][r] = b0[f][r] + sum_o C[f][o][r] * c_src(f)[int2ref[r + off_o]]      (thread per (field, node))
__global__ void __launch_bounds__(128) k_macro_rhs(int fields, int n, int ld, int m, const double* __restrict__ b0,
                                                    const double* __restrict__ C, const int* __restrict__ int2ref,
                                                    const double* __restrict__ c_field0,
                                                    const double* __restrict__ c_field1, int src0, int src1,
                                                    double* __restrict__ b) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= fields * n) return;
  const int f = t / n, r = t % n;
  const double* c = (f == 0 ? src0 : src1) == 0 ? c_field0 : c_field1;
  const double* Cf = C + static_cast<size_t>(f) * NOFF * ld;
  double s = b0[static_cast<size_t>(f) * n + r];
#pragma unroll
  for (int o = 0; o < NOFF; ++o) {
    const int q = r + stencil_offset(o, m);
    if (q >= 0 && q < n) s += Cf[static_cast<size_t>(o) * ld + r] * c[int2ref[q]];
  }
  b[static_cast<size_t>(f) * n + r] = s;
}

// u_all[f][i] = x[f][ref2int[i]] for every macro DoF i; the block [k0, k0+nk) also goes to u / w.
__global__ void __launch_bounds__(128) k_macro_adopt(int fields, int n, const double* __restrict__ x,
                                                      const int* __restrict__ ref2int, double* __restrict__ all,
                                                      size_t all_stride, int k0, int nk, double* __restrict__ u,
                                                      double* __restrict__ w) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= fields * n) return;
  const int f = t / n, i = t % n;
  const double v = x[static_cast<size_t>(f) * n + ref2int[i]];
  all[static_cast<size_t>(f) * all_stride + i] = v;
  if (i >= k0 && i < k0 + nk) (f == 0 ? u : w)[i - k0] = v;
}

// PiSolver::assemble_system (pi_solver.cpp:84-101), the term that depends on the micro solutions:
//   b_i += sum_cells sum_q phi_i(q) rho_q pi_q JxW,   rho_q = sum_a m_a phi_a(q),  pi_q = sum_a g_a phi_a(q),
// QGauss(2) on square cells of side h, with the nodal values (pi_solver.cpp:62-73, :235-246)
//   m = micro mass of the grid at that node (the coupling integral), or min(|m|, 1)            if nonlinear,
//   g = theta * pi_old,                                              or theta * min(|pi|, sqrt|pi|) if nonlinear.
// Thread per macro node (internal ordering r = ix*m + iy): its <= 4 cells in a fixed order, the 4 points of a cell in
// a fixed order, no atomics.  Rows that carry a boundary value keep what the host put there.
__global__ void __launch_bounds__(128) k_macro_product(int n, int m, double h, double theta, int nonlinear,
                                                        const double* __restrict__ c, const int* __restrict__ int2ref,
                                                        const double* __restrict__ x_old,
                                                        const unsigned char* __restrict__ dirichlet,
                                                        double* __restrict__ b) {
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= n || dirichlet[r]) return;
  const int ix = r / m, iy = r - ix * m, nc = m - 1;
  const double gp[2] = {0.5 - 0.28867513459481287, 0.5 + 0.28867513459481287};  // Gauss points of QGauss(2) on [0, 1]
  auto mval = [&](int node) {
    const double v = c[int2ref[node]];
    return nonlinear ? fmin(fabs(v), 1.0) : v;
  };
  auto gval = [&](int node) {
    const double v = x_old[node];
    const double av = fabs(v);
    return nonlinear ? theta * fmin(av, sqrt(av)) : theta * v;
  };
  double sum = 0.0;
  for (int dy = -1; dy <= 0; ++dy)
    for (int dx = -1; dx <= 0; ++dx) {
      const int cx = ix + dx, cy = iy + dy;
      if (cx < 0 || cy < 0 || cx >= nc || cy >= nc) continue;
      const int node[4] = {cx * m + cy, (cx + 1) * m + cy, cx * m + cy + 1, (cx + 1) * m + cy + 1};
      const int mine = (ix - cx) + 2 * (iy - cy);
      double mm[4], gg[4];
      for (int a = 0; a < 4; ++a) {
        mm[a] = mval(node[a]);
        gg[a] = gval(node[a]);
      }
      for (int qj = 0; qj < 2; ++qj)
        for (int qi = 0; qi < 2; ++qi) {
          const double xi = gp[qi], eta = gp[qj];
          const double v[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
          double rho_q = 0.0, pi_q = 0.0;
          for (int a = 0; a < 4; ++a) {
            pi_q += gg[a] * v[a];
            rho_q += mm[a] * v[a];
          }
          sum += v[mine] * (rho_q * pi_q) * (0.25 * h * h);
        }
    }
  b[r] += sum;
}

}  // namespace

namespace msgpu {
void macro_release(msgpu_ctx* ctx) {
  MacroDevice& M = ctx->macro;
  void* ptrs[] = {M.d_diag, M.d_C, M.d_b0, M.d_b, M.d_x, M.d_all, M.d_res, M.d_ref2int, M.d_int2ref, M.d_iters, M.d_flags,
                  M.d_old, M.d_dirichlet};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  M = MacroDevice();
}
}  // namespace msgpu

extern "C" {

int msgpu_macro_setup(msgpu_ctx* ctx, int n_fields, int mesh_kind, int cells_per_axis) {
  if (!ctx || !ctx->setup_done) return MSGPU_ERR_INVALID;
  if (n_fields < 1 || n_fields > 2 || cells_per_axis < 1) return fail(ctx, MSGPU_ERR_INVALID, "macro: 1 or 2 fields");
  const int ntot = ctx->comm_ready ? ctx->n_total : ctx->N;
  const int m = cells_per_axis + 1;
  if (m * m != ntot)
    return fail(ctx, MSGPU_ERR_INVALID, "macro: the macro mesh must have one DoF per grid location (" +
                                            std::to_string(m * m) + " != " + std::to_string(ntot) + ")");
  CK(cudaSetDevice(ctx->device));
  macro_release(ctx);
  MacroDevice& M = ctx->macro;
  try {
    M.ref2int = reference_numbering(mesh_kind, cells_per_axis);
  } catch (const std::exception& ex) {
    return fail(ctx, MSGPU_ERR_INVALID, ex.what());
  }
  M.fields = n_fields;
  M.md.nc = cells_per_axis;
  M.md.m = m;
  M.md.n = m * m;
  M.md.ld = (M.md.n + 1) & ~1;
  M.md.ncells = cells_per_axis * cells_per_axis;
  M.md.q = 0;
  M.md.P = 0;
  M.md.h = 2.0 / cells_per_axis;
  const size_t n = M.md.n, ld = M.md.ld;
  std::vector<int> int2ref(n);
  for (size_t i = 0; i < n; ++i) int2ref[M.ref2int[i]] = static_cast<int>(i);
  CK(cudaMalloc(&M.d_diag, sizeof(double) * n_fields * NDIAG * ld));
  CK(cudaMalloc(&M.d_C, sizeof(double) * n_fields * NOFF * ld));
  CK(cudaMalloc(&M.d_b0, sizeof(double) * n_fields * n));
  CK(cudaMalloc(&M.d_b, sizeof(double) * n_fields * n));
  CK(cudaMalloc(&M.d_x, sizeof(double) * n_fields * n));
  CK(cudaMalloc(&M.d_all, sizeof(double) * n_fields * n));
  CK(cudaMalloc(&M.d_res, sizeof(double) * n_fields));
  CK(cudaMalloc(&M.d_ref2int, sizeof(int) * n));
  CK(cudaMalloc(&M.d_int2ref, sizeof(int) * n));
  CK(cudaMalloc(&M.d_iters, sizeof(int) * n_fields));
  CK(cudaMalloc(&M.d_flags, sizeof(int) * 2));
  CK(cudaMemsetAsync(M.d_diag, 0, sizeof(double) * n_fields * NDIAG * ld, ctx->stream));
  CK(cudaMemsetAsync(M.d_C, 0, sizeof(double) * n_fields * NOFF * ld, ctx->stream));
  CK(cudaMemsetAsync(M.d_b0, 0, sizeof(double) * n_fields * n, ctx->stream));
  CK(cudaMemsetAsync(M.d_x, 0, sizeof(double) * n_fields * n, ctx->stream));
  CK(cudaMemcpyAsync(M.d_ref2int, M.ref2int.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(M.d_int2ref, int2ref.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  M.ready = 1;
  M.source[0] = 0;
  M.source[1] = 1;
  M.rhs_given[0] = M.rhs_given[1] = 0;
  return MSGPU_OK;
}

int msgpu_macro_set_system(msgpu_ctx* ctx, int field, int source, const int* rowstart, const int* col, const double* A,
                           const double* C, const double* b0, const double* x0) {
  if (!ctx || !ctx->macro.ready || field < 0 || field >= ctx->macro.fields || !rowstart || !col || !A)
    return MSGPU_ERR_INVALID;
  if (source < 0 || source > 1) return fail(ctx, MSGPU_ERR_INVALID, "macro: source is 0 (first) or 1 (second coupling vector)");
  CK(cudaSetDevice(ctx->device));
  MacroDevice& M = ctx->macro;
  const int n = M.md.n, m = M.md.m;
  const size_t ld = M.md.ld;
  std::vector<double> diag(NDIAG * ld, 0.0), cop(NOFF * ld, 0.0), v0(n, 0.0), xs(n, 0.0);
  double amax = 0.0;
  for (int k = 0; k < rowstart[n]; ++k) amax = std::fmax(amax, std::fabs(A[k]));
  for (int i = 0; i < n; ++i) {
    const int r = M.ref2int[i];
    for (int k = rowstart[i]; k < rowstart[i + 1]; ++k) {
      const int c = M.ref2int[col[k]];
      // classify by the mesh displacement (unambiguous also for m = 2, where +1 and +(m-1) coincide)
      const int dix = c / m - r / m, diy = c % m - r % m;
      int d = -1, o = -1;  // triangle diagonal, general-operator slot
      if (dix == 0 && diy == 0) d = 0, o = 0;
      else if (dix == 0 && diy == 1) d = 1, o = 1;
      else if (dix == 0 && diy == -1) d = 1, o = 2;
      else if (dix == 1 && diy == -1) d = 2, o = 3;
      else if (dix == -1 && diy == 1) d = 2, o = 4;
      else if (dix == 1 && diy == 0) d = 3, o = 5;
      else if (dix == -1 && diy == 0) d = 3, o = 6;
      else if (dix == 1 && diy == 1) d = 4, o = 7;
      else if (dix == -1 && diy == -1) d = 4, o = 8;
      if (d < 0) {
        if (A[k] != 0.0 || (C && C[k] != 0.0)) return fail(ctx, MSGPU_ERR_INVALID, "macro: entry outside the 9-point stencil");
        continue;
      }
      if (o == 0 || (o & 1)) {
        diag[d * ld + r] = A[k];
      } else {
        // the symmetric partner must agree (the CG kernels store one triangle)
        int kk = -1;
        for (int t = rowstart[col[k]]; t < rowstart[col[k] + 1]; ++t)
          if (col[t] == i) kk = t;
        if (kk < 0 || std::fabs(A[kk] - A[k]) > 1e-12 * amax)
          return fail(ctx, MSGPU_ERR_INVALID, "macro: the system matrix is not symmetric");
      }
      if (C) cop[o * ld + r] = C[k];
    }
    if (b0) v0[r] = b0[i];
    if (x0) xs[r] = x0[i];
  }
  CK(cudaMemcpyAsync(M.d_diag + field * NDIAG * ld, diag.data(), sizeof(double) * NDIAG * ld, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(M.d_C + field * NOFF * ld, cop.data(), sizeof(double) * NOFF * ld, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(M.d_b0 + static_cast<size_t>(field) * n, v0.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (x0)
    CK(cudaMemcpyAsync(M.d_x + static_cast<size_t>(field) * n, xs.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  M.source[field] = source;
  M.has_system[field] = 1;
  return MSGPU_OK;
}

int msgpu_macro_set_rhs(msgpu_ctx* ctx, int field, const double* b) {
  if (!ctx || !ctx->macro.ready || field < 0 || field >= ctx->macro.fields || !b) return MSGPU_ERR_INVALID;
  MacroDevice& M = ctx->macro;
  const int n = M.md.n;
  std::vector<double> v(n);
  for (int i = 0; i < n; ++i) v[M.ref2int[i]] = b[i];
  CK(cudaMemcpyAsync(M.d_b + static_cast<size_t>(field) * n, v.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  M.rhs_given[field] = 1;
  return MSGPU_OK;
}

int msgpu_macro_set_product_load(msgpu_ctx* ctx, int field, double theta, int nonlinear, const double* old0,
                                 const unsigned char* dirichlet) {
  if (!ctx || !ctx->macro.ready || field < 0 || field >= ctx->macro.fields || !old0 || !dirichlet) return MSGPU_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  MacroDevice& M = ctx->macro;
  const int n = M.md.n;
  if (!M.d_old) {
    CK(cudaMalloc(&M.d_old, sizeof(double) * M.fields * n));
    CK(cudaMalloc(&M.d_dirichlet, static_cast<size_t>(M.fields) * n));
  }
  std::vector<double> v(n);
  std::vector<unsigned char> d(n);
  for (int i = 0; i < n; ++i) {
    v[M.ref2int[i]] = old0[i];
    d[M.ref2int[i]] = dirichlet[i];
  }
  CK(cudaMemcpyAsync(M.d_old + static_cast<size_t>(field) * n, v.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(M.d_dirichlet + static_cast<size_t>(field) * n, d.data(), n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  M.product[field] = 1;
  M.theta[field] = theta;
  M.nonlinear[field] = nonlinear != 0;
  return MSGPU_OK;
}

int msgpu_macro_set_solution(msgpu_ctx* ctx, int field, const double* x) {
  if (!ctx || !ctx->macro.ready || field < 0 || field >= ctx->macro.fields || !x) return MSGPU_ERR_INVALID;
  MacroDevice& M = ctx->macro;
  const int n = M.md.n;
  std::vector<double> v(n);
  for (int i = 0; i < n; ++i) v[M.ref2int[i]] = x[i];
  CK(cudaMemcpyAsync(M.d_x + static_cast<size_t>(field) * n, v.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MSGPU_OK;
}

int msgpu_macro_get_solution(msgpu_ctx* ctx, int field, double* x) {
  if (!ctx || !ctx->macro.ready || field < 0 || field >= ctx->macro.fields || !x) return MSGPU_ERR_INVALID;
  MacroDevice& M = ctx->macro;
  const int n = M.md.n;
  std::vector<double> v(n);
  CK(cudaMemcpyAsync(v.data(), M.d_x + static_cast<size_t>(field) * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) x[i] = v[M.ref2int[i]];
  return MSGPU_OK;
}

int msgpu_macro_solve(msgpu_ctx* ctx, double abs_tol, int max_it, int adopt, int* iters) {
  if (!ctx || !ctx->macro.ready) return MSGPU_ERR_INVALID;
  MacroDevice& M = ctx->macro;
  for (int f = 0; f < M.fields; ++f)
    if (!M.has_system[f]) return fail(ctx, MSGPU_ERR_INVALID, "macro: msgpu_macro_set_system was not called for every field");
  CK(cudaSetDevice(ctx->device));
  const int n = M.md.n;
  const size_t cfield = ctx->comm_ready ? static_cast<size_t>(comm_block(ctx)) * ctx->n_ranks : static_cast<size_t>(ctx->N);
  const int blocks = (M.fields * n + 127) / 128;
  // affine right-hand sides from the device-resident coupling vectors (fields with a host-given
  // right-hand side keep it for this one solve)
  if (!(M.rhs_given[0] && (M.fields == 1 || M.rhs_given[1]))) {
    if (M.rhs_given[0] || M.rhs_given[1])
      return fail(ctx, MSGPU_ERR_INVALID, "macro: give the right-hand side of every field or of none");
    if (!ctx->c_view) return fail(ctx, MSGPU_ERR_INVALID, "macro: msgpu_coupling has not been called yet");
    k_macro_rhs<<<blocks, 128, 0, ctx->stream>>>(M.fields, n, M.md.ld, M.md.m, M.d_b0, M.d_C, M.d_int2ref, ctx->c_view,
                                                 ctx->c_view + cfield, M.source[0], M.source[1], M.d_b);
    ++ctx->launches;
  }
  M.rhs_given[0] = M.rhs_given[1] = 0;
  for (int f = 0; f < M.fields; ++f)
    if (M.product[f]) {  // + int phi_i rho_h pi_h from the device-resident micro masses and the previous solution
      if (!ctx->c_view) return fail(ctx, MSGPU_ERR_INVALID, "macro: msgpu_coupling has not been called yet");
      k_macro_product<<<(n + 127) / 128, 128, 0, ctx->stream>>>(
          n, M.md.m, M.md.h, M.theta[f], M.nonlinear[f], M.source[f] == 0 ? ctx->c_view : ctx->c_view + cfield, M.d_int2ref,
          M.d_old + static_cast<size_t>(f) * n, M.d_dirichlet + static_cast<size_t>(f) * n, M.d_b + static_cast<size_t>(f) * n);
      ++ctx->launches;
    }
  CgArgs a;
  a.md = M.md;
  a.N = M.fields;
  a.diag = M.d_diag;
  a.diag_stride = static_cast<long long>(NDIAG) * M.md.ld;
  a.b = M.d_b;
  a.x = M.d_x;
  a.tol = abs_tol;
  a.max_it = max_it;
  a.precond = 0;
  a.omega = 1.0;
  a.iters = M.d_iters;
  a.res = M.d_res;
  a.work_counter = M.d_flags;
  a.fail_flag = M.d_flags + 1;
  a.sw = &ctx->cg_sw;
  a.workspace = &ctx->d_cg_ws;
  a.workspace_bytes = &ctx->cg_ws_bytes;
  int variant = 0;
  const int launched = launch_cg(ctx->stream, a, ctx->num_sms, &variant);
  if (launched < 0) return fail(ctx, MSGPU_ERR_CUDA, "macro: the CG kernel could not be launched");
  ctx->launches += launched;
  for (int f = 0; f < M.fields; ++f)
    if (M.product[f])  // old_solution = solution (pi_solver.cpp:259, :264)
      CK(cudaMemcpyAsync(M.d_old + static_cast<size_t>(f) * n, M.d_x + static_cast<size_t>(f) * n, sizeof(double) * n,
                         cudaMemcpyDeviceToDevice, ctx->stream));
  if (adopt) {
    k_macro_adopt<<<blocks, 128, 0, ctx->stream>>>(M.fields, n, M.d_x, M.d_ref2int, M.d_all, static_cast<size_t>(n),
                                                   ctx->comm_ready ? ctx->k_offset : 0, ctx->N, ctx->d_u, ctx->d_w);
    ++ctx->launches;
  }
  int flags[2] = {0, 0}, its[2] = {0, 0};
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(flags, M.d_flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(its, M.d_iters, sizeof(int) * M.fields, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (iters)
    for (int f = 0; f < M.fields; ++f) iters[f] = its[f];
  if (flags[1]) return fail(ctx, MSGPU_ERR_NO_CONVERGENCE, "SolverControl::NoConvergence in the macro system");
  return MSGPU_OK;
}

}  // extern "C"
