// Per-grid error / residual norms (K8) and the elliptic-parabolic micro solver pieces (K7).
//
// Replaces
//   MicroSolver::compute_all_errors per-grid part   src/elliptic-elliptic/micro.cpp:189-205,
//                                                   src/bio-multiscale/micro.cpp:323-338
//   MicroSolver::get_residual                       src/bio-multiscale/micro.cpp:367-378
//   RhoSolver::setup_system / setup_scatter / assemble_system / compute_error
//                                                   src/elliptic-parabolic/rho_solver.cpp:55-93, :111-218, :307-330
#include <algorithm>
#include <cmath>
#include <vector>

#include "ctx.h"
#include "extra_bodies.cuh"
#include "fe_tables.h"

namespace msgpu {

namespace {

constexpr int TBX = 128;

__global__ void __launch_bounds__(TBX) k_cell_errors(const __grid_constant__ VmProgram prog, MeshDims md,
                                                      const double* __restrict__ slots, int n_slots, int qe,
                                                      const double* __restrict__ pt, const double* __restrict__ wt,
                                                      const double* __restrict__ x, int N, double fd,
                                                      double* __restrict__ out) {
  extern __shared__ double R[];
  body_cell_errors(prog, md, slots, n_slots, qe, pt, wt, x, N, fd, out,
                   static_cast<long long>(blockIdx.x) * TBX + threadIdx.x, R + threadIdx.x, TBX);
}

__global__ void __launch_bounds__(TBX) k_cell_residuals(MeshDims md, const double* __restrict__ x,
                                                         const double* __restrict__ xold, int N,
                                                         double* __restrict__ out) {
  body_cell_residual(md, x, xold, N, out, static_cast<long long>(blockIdx.x) * TBX + threadIdx.x);
}

__global__ void __launch_bounds__(TBX) k_cell_point_load(const __grid_constant__ VmProgram prog, MeshDims md,
                                                          const double* __restrict__ slots, int n_slots, int qe,
                                                          const double* __restrict__ pt, const double* __restrict__ wt,
                                                          int k0, int nk, double* __restrict__ cell) {
  extern __shared__ double R[];
  body_cell_point_load(prog, md, slots, n_slots, qe, pt, wt, k0, nk, cell,
                       static_cast<long long>(blockIdx.x) * TBX + threadIdx.x, R + threadIdx.x, TBX);
}

// One warp per (system, quantity): fixed-order sum over the cells (segmented warp reduction).
__global__ void __launch_bounds__(TBX) k_reduce_cells(const double* __restrict__ cellred, int N, int ncells,
                                                       double* __restrict__ out2) {
  const int warp = (blockIdx.x * TBX + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * N) return;
  const double* v = cellred + static_cast<size_t>(warp) * ncells;
  double s = 0.0;
  for (int c = lane; c < ncells; c += 32) s += v[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out2[warp] = s;
}

inline int blocks(long long work) { return static_cast<int>((work + TBX - 1) / TBX); }
inline size_t smem_for(const VmProgram& p) { return static_cast<size_t>(std::max(8, p.n_regs)) * TBX * sizeof(double); }

#define CKVM()                                                              \
  do {                                                                      \
    if (const char* e_ = vm_launch_error()) {                               \
      ctx->err = e_;                                                        \
      vm_clear_launch_error();                                              \
      return MSGPU_ERR_UNSUPPORTED;                                         \
    }                                                                       \
  } while (0)

#define CKX(call)                                                           \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);        \
      return MSGPU_ERR_CUDA;                                                \
    }                                                                       \
  } while (0)

template <typename T>
int ensure(msgpu_ctx* ctx, T** p, size_t count) {
  if (*p) return MSGPU_OK;
  CKX(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return MSGPU_OK;
}

}  // namespace

int launch_cell_errors(cudaStream_t s, const VmProgram& prog, MeshDims md, const double* slots, int n_slots, int qe,
                       const double* pt, const double* wt, const double* x, int N, double fd, double* cellred,
                       double* out2) {
  const long long work = static_cast<long long>(N) * md.ncells;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_cell_errors), smem_for(prog), "the exact solution")) return 0;
  k_cell_errors<<<blocks(work), TBX, smem_for(prog), s>>>(prog, md, slots, n_slots, qe, pt, wt, x, N, fd, cellred);
  k_reduce_cells<<<blocks(static_cast<long long>(2) * N * 32), TBX, 0, s>>>(cellred, N, md.ncells, out2);
  return 2;
}

static int launch_reduce_cells(cudaStream_t s, const double* cellred, int N, int ncells, double* out2) {
  k_reduce_cells<<<blocks(static_cast<long long>(2) * N * 32), TBX, 0, s>>>(cellred, N, ncells, out2);
  return 1;
}

int launch_cell_residuals(cudaStream_t s, MeshDims md, const double* x, const double* xold, int N, double* cellred,
                          double* out2) {
  const long long work = static_cast<long long>(N) * md.ncells;
  k_cell_residuals<<<blocks(work), TBX, 0, s>>>(md, x, xold, N, cellred);
  k_reduce_cells<<<blocks(static_cast<long long>(2) * N * 32), TBX, 0, s>>>(cellred, N, md.ncells, out2);
  return 2;
}

int launch_cell_point_load(cudaStream_t s, const VmProgram& prog, MeshDims md, const double* slots, int n_slots, int qe,
                           const double* pt, const double* wt, int k0, int nk, double* cell) {
  const long long work = static_cast<long long>(nk) * md.ncells;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_cell_point_load), smem_for(prog), "the projected function")) return 0;
  k_cell_point_load<<<blocks(work), TBX, smem_for(prog), s>>>(prog, md, slots, n_slots, qe, pt, wt, k0, nk, cell);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// errors / residuals
static int error_quadrature(msgpu_ctx* ctx) {
  // elliptic / bio: QGauss(fem_q_deg = 12) (micro.cpp:194); parabolic: QGauss(3) (rho_solver.cpp:316)
  const int qe = ctx->desc.problem == MSGPU_PARABOLIC ? 3 : ctx->desc.q_degree;
  if (ctx->d_eq_pt && ctx->eq_degree == qe) return MSGPU_OK;
  std::vector<double> pt, wt;
  gauss_legendre_unit(qe, pt, wt);
  int rc;
  if ((rc = ensure(ctx, &ctx->d_eq_pt, 16))) return rc;
  if ((rc = ensure(ctx, &ctx->d_eq_wt, 16))) return rc;
  CKX(cudaMemcpy(ctx->d_eq_pt, pt.data(), sizeof(double) * qe, cudaMemcpyHostToDevice));
  CKX(cudaMemcpy(ctx->d_eq_wt, wt.data(), sizeof(double) * qe, cudaMemcpyHostToDevice));
  ctx->eq_degree = qe;
  return MSGPU_OK;
}

// d_err[2k], d_err[2k+1] = the per-grid sums of squares (L2 and H1-seminorm parts); no copy to the host
int launch_grid_error_sums(msgpu_ctx* ctx) {
  if (ctx->pp.prog_err < 0) {
    ctx->err = "no exact solution was set (msgpu_set_function(MSGPU_FN_SOLUTION, ...))";
    return MSGPU_ERR_INVALID;
  }
  CKX(cudaSetDevice(ctx->device));
  int rc = error_quadrature(ctx);
  if (rc) return rc;
  const MeshDims& md = ctx->md;
  const int N = ctx->N;
  if ((rc = ensure(ctx, &ctx->d_cellred, static_cast<size_t>(N) * 2 * md.ncells))) return rc;
  // the exact-solution program specialised at first use (jit.h); the interpreter kernel is the fallback
  const VmProgram& exact = ctx->pp.ps.point_program(ctx->pp.prog_err);
  if (ctx->jit_errors_ready == 0) {
    const char* jit_env = std::getenv("MSGPU_JIT");
    ctx->jit_errors_ready = -1;
    if (!(jit_env && jit_env[0] == '0')) {
      const std::string src = cell_errors_source(exact);
      std::string log;
      if (!src.empty() && jit_compile(src, "k_cell_errors_jit", &ctx->jit_errors, &log) == 0) {
        ctx->jit_errors_ready = 1;
        ctx->jit_kernels += 1;
      } else if (ctx->jit_log.empty()) {
        ctx->jit_log = src.empty() ? "the exact-solution program uses an instruction the code generator does not lower" : log;
      }
    }
  }
  int launched = -1;
  if (ctx->jit_errors_ready == 1)
    launched = launch_cell_errors_jit(ctx->stream, ctx->jit_errors, md, ctx->d_slots, ctx->pp.ps.n_slots(),
                                      ctx->eq_degree, ctx->d_eq_pt, ctx->d_eq_wt, ctx->d_x, N, 1e-8, ctx->d_cellred);
  if (launched > 0)
    ctx->launches += launched + launch_reduce_cells(ctx->stream, ctx->d_cellred, N, md.ncells, ctx->d_err);
  else
    ctx->launches += launch_cell_errors(ctx->stream, exact, md, ctx->d_slots, ctx->pp.ps.n_slots(), ctx->eq_degree,
                                        ctx->d_eq_pt, ctx->d_eq_wt, ctx->d_x, N, 1e-8, ctx->d_cellred, ctx->d_err);
  CKVM();
  CKX(cudaGetLastError());
  return MSGPU_OK;
}

int compute_errors(msgpu_ctx* ctx, double* l2, double* h1) {
  const int rc0 = launch_grid_error_sums(ctx);
  if (rc0) return rc0;
  const int N = ctx->N;
  std::vector<double> sums(2 * static_cast<size_t>(N));
  CKX(cudaMemcpyAsync(sums.data(), ctx->d_err, sizeof(double) * sums.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CKX(cudaStreamSynchronize(ctx->stream));
  const bool full_h1 = ctx->desc.problem == MSGPU_PARABOLIC;  // H1_norm per cell there (rho_solver.cpp:319-323)
  for (int k = 0; k < N; ++k) {
    l2[k] = std::sqrt(sums[2 * k]);
    if (h1) h1[k] = std::sqrt(full_h1 ? sums[2 * k] + sums[2 * k + 1] : sums[2 * k + 1]);
  }
  return MSGPU_OK;
}

// d_err[2k] = ||x_k - x_k_old||^2 per grid; no copy to the host
int launch_grid_residual_sums(msgpu_ctx* ctx) {
  if (!ctx->d_xold) {
    ctx->err = "residuals are defined for the bio and parabolic problems";
    return MSGPU_ERR_INVALID;
  }
  CKX(cudaSetDevice(ctx->device));
  const MeshDims& md = ctx->md;
  const int N = ctx->N;
  int rc;
  if ((rc = ensure(ctx, &ctx->d_cellred, static_cast<size_t>(N) * 2 * md.ncells))) return rc;
  ctx->launches += launch_cell_residuals(ctx->stream, md, ctx->d_x, ctx->d_xold, N, ctx->d_cellred, ctx->d_err);
  CKX(cudaGetLastError());
  return MSGPU_OK;
}

int compute_residuals(msgpu_ctx* ctx, double* l2) {
  const int rc0 = launch_grid_residual_sums(ctx);
  if (rc0) return rc0;
  const int N = ctx->N;
  std::vector<double> sums(2 * static_cast<size_t>(N));
  CKX(cudaMemcpyAsync(sums.data(), ctx->d_err, sizeof(double) * sums.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CKX(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < N; ++k) l2[k] = std::sqrt(sums[2 * k]);
  return MSGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// Batched L2 projection onto the Q1 space of every grid: VectorTools::project(dof_handler, constraints,
// QGauss(qe), f, solutions[k]) for all k (rho_solver.cpp:80-83 for init_rho; micro.cpp:229-233,
// bio micro.cpp:386-390, rho_solver.cpp:355-358 for the exact solution).  Load vectors by
// k_cell_point_load + k_gather, then M x = load with the batched CG on the one shared mass stencil.
// Overwrites the solutions AND the right-hand sides of the context.
int project_into_solutions(msgpu_ctx* ctx, const VmProgram& prog, int qe, const char* what) {
  CKX(cudaSetDevice(ctx->device));
  const MeshDims& md = ctx->md;
  const int N = ctx->N;
  int rc;
  if (!ctx->d_mass) {  // elliptic / bio contexts: the mass stencil is only needed here
    if ((rc = ensure(ctx, &ctx->d_mass, static_cast<size_t>(NDIAG) * md.ld))) return rc;
    std::vector<double> mass, sys, lumped;
    build_stencils(md, 0.0, 0.0, 0.0, 0.0, mass, sys, lumped);
    CKX(cudaMemcpy(ctx->d_mass, mass.data(), sizeof(double) * mass.size(), cudaMemcpyHostToDevice));
  }
  std::vector<double> pt, wt;
  gauss_legendre_unit(qe, pt, wt);
  struct DeviceArray {  // freed on every return path
    double* p = nullptr;
    ~DeviceArray() { if (p) cudaFree(p); }
  } q_pt, q_wt;
  CKX(cudaMalloc(reinterpret_cast<void**>(&q_pt.p), qe * sizeof(double)));
  CKX(cudaMalloc(reinterpret_cast<void**>(&q_wt.p), qe * sizeof(double)));
  double *const d_pt = q_pt.p, *const d_wt = q_wt.p;
  CKX(cudaMemcpy(d_pt, pt.data(), qe * sizeof(double), cudaMemcpyHostToDevice));
  CKX(cudaMemcpy(d_wt, wt.data(), qe * sizeof(double), cudaMemcpyHostToDevice));
  for (int k0 = 0; k0 < N; k0 += ctx->chunk) {
    const int nk = std::min(ctx->chunk, N - k0);
    ctx->launches += launch_cell_point_load(ctx->stream, prog, md, ctx->d_slots, ctx->pp.ps.n_slots(), qe, d_pt, d_wt,
                                            k0, nk, ctx->d_cell);
    GatherArgs g{};
    g.md = md;
    g.problem = PF_PARABOLIC;
    g.k0 = k0;
    g.nk = nk;
    g.cell = ctx->d_cell;
    g.face = nullptr;
    g.diag = nullptr;
    g.b0 = ctx->d_b;  // load vector straight into the right-hand side of the mass solve
    g.jw = nullptr;
    g.edge = ctx->d_edge;
    ctx->launches += launch_gather(ctx->stream, g);
  }
  CKVM();
  CKX(cudaMemsetAsync(ctx->d_x, 0, sizeof(double) * N * md.n, ctx->stream));
  CgArgs a{};
  a.md = md;
  a.N = N;
  a.diag = ctx->d_mass;
  a.diag_stride = 0;
  a.b = ctx->d_b;
  a.x = ctx->d_x;
  a.tol = 1e-14;
  a.max_it = 10 * md.n + 100;
  a.precond = 0;
  a.iters = ctx->d_iters;
  a.res = ctx->d_res;
  a.work_counter = ctx->d_flags + FLAG_COUNTER;
  a.fail_flag = ctx->d_flags + FLAG_CGFAIL;
  a.sw = &ctx->cg_sw;
  a.workspace = &ctx->d_cg_ws;
  a.workspace_bytes = &ctx->cg_ws_bytes;
  int variant = 0;
  const int l = launch_cg(ctx->stream, a, ctx->num_sms, &variant);
  if (l < 0) {
    ctx->err = "cannot allocate the CG workspace";
    return MSGPU_ERR_CUDA;
  }
  ctx->launches += l;
  int flags[4];
  CKX(cudaMemcpyAsync(flags, ctx->d_flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
  CKX(cudaStreamSynchronize(ctx->stream));
  if (flags[FLAG_CGFAIL]) {
    ctx->err = std::string("L2 projection of ") + what + " did not converge";
    return MSGPU_ERR_NO_CONVERGENCE;
  }
  return MSGPU_OK;
}

int build_tables_at(msgpu_ctx* ctx, double t);

// MicroSolver::set_exact_solution (micro.cpp:224-234, bio micro.cpp:381-391, rho_solver.cpp:353-359)
int set_exact_solution(msgpu_ctx* ctx, double t) {
  if (ctx->pp.prog_err < 0) {
    ctx->err = "no exact solution was set (msgpu_set_function(MSGPU_FN_SOLUTION, ...))";
    return MSGPU_ERR_INVALID;
  }
  int qe = ctx->desc.q_degree;  // QGauss(fem_q_deg) in the elliptic and bio drivers
  if (ctx->desc.problem == MSGPU_PARABOLIC) {
    qe = 3;  // rho_solver.cpp:357
    const int rc = build_tables_at(ctx, t);
    if (rc) return rc;
  }
  return project_into_solutions(ctx, ctx->pp.ps.point_program(ctx->pp.prog_err), qe, "the exact solution");
}

// ------------------------------------------------------------------------------------------------
// parabolic: shared stencils built on the host (n entries, closed-form Q1 local matrices on a
// square cell: what MatrixCreator::create_mass_matrix / create_laplace_matrix with QGauss(2)
// integrate exactly, rho_solver.cpp:64-65)
int parabolic_setup(msgpu_ctx* ctx) {
  CKX(cudaSetDevice(ctx->device));
  const MeshDims& md = ctx->md;
  const int N = ctx->N;
  int rc;
  if ((rc = ensure(ctx, &ctx->d_mass, static_cast<size_t>(NDIAG) * md.ld))) return rc;
  if ((rc = ensure(ctx, &ctx->d_lumped, static_cast<size_t>(md.n)))) return rc;
  std::vector<double> mass, sys, lumped;
  build_stencils(md, 0.0, ctx->scalars[3], ctx->scalars[0], ctx->scalars[2], mass, sys, lumped);
  CKX(cudaMemcpy(ctx->d_mass, mass.data(), sizeof(double) * mass.size(), cudaMemcpyHostToDevice));
  CKX(cudaMemcpy(ctx->d_lumped, lumped.data(), sizeof(double) * lumped.size(), cudaMemcpyHostToDevice));
  ctx->parabolic_dt = -1.0;
  if (ctx->pp.prog_init < 0) return MSGPU_OK;  // start from zero
  // rho_k^0 = L2 projection of init_rho(x_k, .) with QGauss(3) (rho_solver.cpp:78-83)
  if ((rc = project_into_solutions(ctx, ctx->pp.ps.point_program(ctx->pp.prog_init), 3, "init_rho"))) return rc;
  CKX(cudaMemcpyAsync(ctx->d_xold, ctx->d_x, sizeof(double) * N * md.n, cudaMemcpyDeviceToDevice, ctx->stream));
  CKX(cudaStreamSynchronize(ctx->stream));
  return MSGPU_OK;
}

// K0 tables for time t (declared in msgpu.cu)
int build_tables_at(msgpu_ctx* ctx, double t);

int parabolic_assemble(msgpu_ctx* ctx, double dt, double t) {
  CKX(cudaSetDevice(ctx->device));
  const MeshDims& md = ctx->md;
  const ProgramSet& ps = ctx->pp.ps;
  const int N = ctx->N;
  const double kappa = ctx->scalars[0], p_F = ctx->scalars[1], R = ctx->scalars[2], D = ctx->scalars[3];
  if (dt != ctx->parabolic_dt) {  // A = M + dt D L + kappa dt R (Robin mass): identical for every grid
    std::vector<double> mass, sys, lumped;
    build_stencils(md, dt, D, kappa, R, mass, sys, lumped);
    CKX(cudaMemcpyAsync(ctx->d_diag, sys.data(), sizeof(double) * sys.size(), cudaMemcpyHostToDevice, ctx->stream));
    CKX(cudaStreamSynchronize(ctx->stream));
    ctx->parabolic_dt = dt;
    ctx->stencil_hint = 0;  // another matrix
  }
  int rc = build_tables_at(ctx, t);  // data.set_time(t) (time_manager.cpp:66)
  if (rc) return rc;
  Tables tb;
  tb.slots = ctx->d_slots;
  tb.T0 = ctx->d_T0;
  tb.T1 = ctx->d_T1;
  tb.n_slots = ps.n_slots();
  tb.n0 = ps.n_line0();
  tb.n1 = ps.n_line1();
  MomentPrograms mp;
  mp.full = mp.F = mp.G = &ps.point_program(ctx->pp.prog_vol_G);
  mp.has_map = 0;
  mp.jac_depends_on_y = 0;
  for (int k0 = 0; k0 < N; k0 += ctx->chunk) {
    const int nk = std::min(ctx->chunk, N - k0);
    int launched = -1;
    if (ctx->jit_kernels)
      launched = launch_cell_moments_jit(ctx->stream, ctx->jit_moments, md, tb, k0, nk, 1.0, ctx->d_cell,
                                         ctx->d_flags + FLAG_JACOBIAN);
    if (launched < 0)
      launched = launch_cell_moments(ctx->stream, mp, md, tb, ctx->d_qtab,
                                     ctx->d_qtab + static_cast<size_t>(md.q) * md.q * QTAB_STRIDE, k0, nk, 1.0,
                                     ctx->d_cell, ctx->d_flags + FLAG_JACOBIAN);
    ctx->launches += launched;
    const VmProgram* faces[4];
    for (int s = 0; s < 4; ++s) faces[s] = &ps.point_program(ctx->pp.prog_face[s]);
    ctx->launches += launch_face_moments_all(ctx->stream, faces, md, tb, ctx->d_q1pt, ctx->d_q1wt, k0, nk, 0, ctx->d_face);
    GatherArgs g{};
    g.md = md;
    g.problem = PF_PARABOLIC;
    g.k0 = k0;
    g.nk = nk;
    g.cell = ctx->d_cell;
    g.face = ctx->d_face;
    g.robin_left = g.robin_right = 0.0;
    g.diag = nullptr;
    g.b0 = ctx->d_b0;
    g.jw = nullptr;
    g.edge = ctx->d_edge;
    ctx->launches += launch_gather(ctx->stream, g);
  }
  ctx->launches += launch_shared_spmv(ctx->stream, md, N, ctx->d_mass, ctx->d_xold, ctx->d_b);
  RhsArgs r{};
  r.md = md;
  r.problem = PF_PARABOLIC;
  r.N = N;
  r.u = ctx->d_u;
  r.w = ctx->d_w;
  r.coef_u = kappa;
  r.coef_w = p_F;
  r.dt = dt;
  r.diag = ctx->d_diag;
  r.b0 = ctx->d_b0;
  r.jw = nullptr;
  r.edge = ctx->d_edge;
  r.bv = nullptr;
  r.b = ctx->d_b;
  r.x = ctx->d_x;
  ctx->launches += launch_rhs_and_dirichlet(ctx->stream, r);
  CKVM();
  CKX(cudaGetLastError());
  return MSGPU_OK;
}

}  // namespace msgpu
