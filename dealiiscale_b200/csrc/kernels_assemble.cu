// Batched micro assembly on sm_100a: tables (K0), cell / face moments (K1), stencil gather,
// Dirichlet elimination and right-hand sides (K2).
//
// Replaces, for all N micro systems at once,
//   MicroSolver::compute_pullback_objects + integrate_cell + assemble_system
//       (reference: src/elliptic-elliptic/micro.cpp:75-171)
//   MicroSolver::local_assemble_system + copy_local_to_global
//       (reference: src/bio-multiscale/micro.cpp:183-288)
// The work is FP64-ALU bound (SURVEY.md §7 "Assembly is FP64-compute-heavy"); memory traffic is
// the coalesced write of 18 doubles per cell and the read of the line tables, which are
// warp-uniform (y0 side) or unit-stride (y1 side) by construction of the device ordering.
#include <cstdio>
#include <cstdlib>

#include "assemble_bodies.cuh"
#include "kernels.h"

namespace msgpu {

namespace {

constexpr int TB = 128;  // threads per block of the assembly kernels

inline size_t vm_smem(const VmProgram& p, int threads) {
  int regs = p.n_regs < 8 ? 8 : p.n_regs;
  return static_cast<size_t>(regs) * threads * sizeof(double);
}

inline int blocks_for(long long work) { return static_cast<int>((work + TB - 1) / TB); }

// Per-quadrature-point basis products in constant memory: every lane of a warp reads the same
// entry, so these become uniform constant-bank operands.  The contents depend on q only.
__constant__ double c_qtab12[144 * QTAB_STRIDE];
__constant__ double c_qsum12[QTAB_STRIDE];
__constant__ double c_qtab3[9 * QTAB_STRIDE];
__constant__ double c_qsum3[QTAB_STRIDE];
__constant__ double c_qtab2[4 * QTAB_STRIDE];
__constant__ double c_qsum2[QTAB_STRIDE];
__constant__ double c_q1d12[Q1D_ROWS * 12];  // fe_tables.h build_q1d
__constant__ double c_q1d3[Q1D_ROWS * 3];
__constant__ double c_q1d2[Q1D_ROWS * 2];

template <int Q>
struct TabConst;
#ifdef __CUDA_ARCH__
#define MSGPU_CONST_TAB(Q, TAB, SUM, LINE)                                                  \
  template <>                                                                               \
  struct TabConst<Q> {                                                                      \
    __host__ __device__ double operator()(int q, int i) const { return TAB[q * QTAB_STRIDE + i]; } \
    __host__ __device__ double total(int i) const { return SUM[i]; }                        \
    __host__ __device__ double line(int r, int i) const { return LINE[r * Q + i]; }         \
  };
#else
#define MSGPU_CONST_TAB(Q, TAB, SUM, LINE)                                                  \
  template <>                                                                               \
  struct TabConst<Q> {                                                                      \
    __host__ __device__ double operator()(int, int) const { return 0.0; }                   \
    __host__ __device__ double total(int) const { return 0.0; }                             \
    __host__ __device__ double line(int, int) const { return 0.0; }                         \
  };
#endif
MSGPU_CONST_TAB(12, c_qtab12, c_qsum12, c_q1d12)
MSGPU_CONST_TAB(3, c_qtab3, c_qsum3, c_q1d3)
MSGPU_CONST_TAB(2, c_qtab2, c_qsum2, c_q1d2)

template <int Q, int W, int MODE>
__global__ void __launch_bounds__(TB) k_cell_moments_vec(const __grid_constant__ VmProgram progF,
                                                          const __grid_constant__ VmProgram prog, MeshDims md, Tables tb,
                                                          int k0, int nk, double stiffness_scale,
                                                          double* __restrict__ cell, int* __restrict__ error_flag) {
  extern __shared__ double R[];
  body_cell_moments_vec<Q, W, MODE>(progF, prog, md, tb, TabConst<Q>(), k0, nk, stiffness_scale, cell, error_flag,
                                    static_cast<long long>(blockIdx.x) * TB + threadIdx.x, R + threadIdx.x, TB);
}

template <int Q, int W, int MODE>
int launch_moments_vec(cudaStream_t s, const VmProgram& progF, const VmProgram& prog, MeshDims md, Tables tb, int k0,
                       int nk, double scale, double* cell, int* error_flag) {
  size_t regs = static_cast<size_t>(prog.n_regs) * W;
  if (MODE == MODE_CONST_F && static_cast<size_t>(progF.n_regs) > regs) regs = progF.n_regs;
  if (regs < 8) regs = 8;
  const size_t smem = regs * TB * sizeof(double);
  if (smem > 200 * 1024) return -1;  // the caller falls back to one VM run per point (W times fewer registers)
  auto kern = k_cell_moments_vec<Q, W, MODE>;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(kern), smem, "the volume integrand")) return 0;
  kern<<<blocks_for(static_cast<long long>(nk) * md.ncells), TB, smem, s>>>(progF, prog, md, tb, k0, nk, scale, cell,
                                                                            error_flag);
  return 1;
}

template <int Q, int WG, int WC>
int dispatch_moments(cudaStream_t s, const MomentPrograms& mp, MeshDims md, Tables tb, int k0, int nk, double scale,
                     double* cell, int* error_flag) {
  if (!mp.has_map) return launch_moments_vec<Q, WC, MODE_NO_MAP>(s, *mp.G, *mp.G, md, tb, k0, nk, scale, cell, error_flag);
  if (!mp.jac_depends_on_y)
    return launch_moments_vec<Q, WC, MODE_CONST_F>(s, *mp.F, *mp.G, md, tb, k0, nk, scale, cell, error_flag);
  return launch_moments_vec<Q, WG, MODE_GENERAL>(s, *mp.full, *mp.full, md, tb, k0, nk, scale, cell, error_flag);
}

// Thin __global__ wrappers: thread id -> body.  TB threads per block, one VM register file of
// n_regs doubles per thread in dynamic shared memory (register j of thread t at R[j*TB + t]).
__global__ void __launch_bounds__(TB) k_slots(const __grid_constant__ VmProgram prog, const double* __restrict__ xy,
                                               int N, double t, double* __restrict__ slots, int n_slots) {
  extern __shared__ double R[];
  body_slots(prog, xy, N, t, slots, n_slots, static_cast<long long>(blockIdx.x) * TB + threadIdx.x, R + threadIdx.x, TB);
}
__global__ void __launch_bounds__(TB) k_lines(const __grid_constant__ VmProgram prog0,
                                               const __grid_constant__ VmProgram prog1,
                                               const double* __restrict__ coord0, const double* __restrict__ coord1,
                                               int P, int N, const double* __restrict__ slots, int n_slots,
                                               double* __restrict__ T0, int n0, double* __restrict__ T1, int n1) {
  extern __shared__ double R[];
  body_lines(prog0, prog1, coord0, coord1, P, N, slots, n_slots, T0, n0, T1, n1,
             static_cast<long long>(blockIdx.x) * TB + threadIdx.x, R + threadIdx.x, TB);
}
__global__ void __launch_bounds__(TB) k_cell_moments(const __grid_constant__ VmProgram prog, MeshDims md, Tables tb,
                                                      const double* __restrict__ qtab, int k0, int nk,
                                                      double stiffness_scale, int has_map,
                                                      double* __restrict__ cell, int* __restrict__ error_flag) {
  extern __shared__ double R[];
  body_cell_moments(prog, md, tb, qtab, k0, nk, stiffness_scale, has_map, cell, error_flag,
                    static_cast<long long>(blockIdx.x) * TB + threadIdx.x, R + threadIdx.x, TB);
}
__global__ void __launch_bounds__(TB) k_face_moments(const __grid_constant__ VmProgram prog, int side, MeshDims md,
                                                      Tables tb, const double* __restrict__ q1d_pt,
                                                      const double* __restrict__ q1d_wt, int k0, int nk, int has_map,
                                                      double* __restrict__ face) {
  extern __shared__ double R[];
  body_face_moments(prog, side, md, tb, q1d_pt, q1d_wt, k0, nk, has_map, face,
                    static_cast<long long>(blockIdx.x) * TB + threadIdx.x, R + threadIdx.x, TB);
}
// all four sides in one launch (blockIdx.y = side): the four launches were latency-bound
__global__ void __launch_bounds__(TB) k_face_moments4(const __grid_constant__ VmProgram p0,
                                                       const __grid_constant__ VmProgram p1,
                                                       const __grid_constant__ VmProgram p2,
                                                       const __grid_constant__ VmProgram p3, MeshDims md, Tables tb,
                                                       const double* __restrict__ q1d_pt,
                                                       const double* __restrict__ q1d_wt, int k0, int nk, int has_map,
                                                       double* __restrict__ face) {
  extern __shared__ double R[];
  const int side = blockIdx.y;
  const VmProgram& prog = side == 0 ? p0 : side == 1 ? p1 : side == 2 ? p2 : p3;
  body_face_moments(prog, side, md, tb, q1d_pt, q1d_wt, k0, nk, has_map, face,
                    static_cast<long long>(blockIdx.x) * TB + threadIdx.x, R + threadIdx.x, TB);
}
// blockIdx.x: system of the chunk, blockIdx.y: block of TB nodes (no division of a 64-bit thread index)
template <bool UNIFORM>
__global__ void __launch_bounds__(TB) k_gather(const __grid_constant__ GatherArgs a) {
  const int r = static_cast<int>(blockIdx.y) * TB + static_cast<int>(threadIdx.x);
  if (r < a.md.n) body_gather_node<UNIFORM>(a, static_cast<int>(blockIdx.x), r);
}
__global__ void __launch_bounds__(TB) k_boundary_values(const __grid_constant__ VmProgram prog, MeshDims md, Tables tb,
                                                         int N, double* __restrict__ bv) {
  extern __shared__ double R[];
  body_boundary_values(prog, md, tb, N, bv, static_cast<long long>(blockIdx.x) * TB + threadIdx.x, R + threadIdx.x,
                       TB);
}
__global__ void __launch_bounds__(TB) k_rhs(RhsArgs a) {
  const int r = static_cast<int>(blockIdx.y) * TB + static_cast<int>(threadIdx.x);
  if (r < a.md.n) body_rhs_node(a, static_cast<int>(blockIdx.x), r);
}
__global__ void __launch_bounds__(TB) k_mask_dirichlet(MeshDims md, int N, double* __restrict__ diag) {
  const int r = static_cast<int>(blockIdx.y) * TB + static_cast<int>(threadIdx.x);
  if (r < md.n && static_cast<int>(blockIdx.x) < N) body_mask_dirichlet_node(md, diag, static_cast<int>(blockIdx.x), r);
}

thread_local char g_vm_error[256] = {0};

}  // namespace

bool vm_prepare_launch(const void* kernel, size_t smem_bytes, const char* what) {
  constexpr size_t kLimit = 220 * 1024;
  if (smem_bytes > kLimit) {
    std::snprintf(g_vm_error, sizeof g_vm_error,
                  "%s needs %zu live temporaries per evaluation point; the interpreter kernels hold at most %zu (%zu KB of shared "
                  "memory per SM)",
                  what, smem_bytes / (TB * sizeof(double)), kLimit / (TB * sizeof(double)), kLimit / 1024);
    return false;
  }
  if (smem_bytes > 48 * 1024 &&
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)) != cudaSuccess) {
    cudaGetLastError();
    std::snprintf(g_vm_error, sizeof g_vm_error, "%s: cannot opt in to %zu bytes of dynamic shared memory", what, smem_bytes);
    return false;
  }
  return true;
}
const char* vm_launch_error() { return g_vm_error[0] ? g_vm_error : nullptr; }
void vm_clear_launch_error() { g_vm_error[0] = 0; }

int launch_slots(cudaStream_t s, const VmProgram& prog, const double* xy, int N, double t, double* slots, int n_slots) {
  if (n_slots == 0 || N == 0) return 0;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_slots), vm_smem(prog, TB), "the per-system part of the expressions"))
    return 0;
  k_slots<<<blocks_for(N), TB, vm_smem(prog, TB), s>>>(prog, xy, N, t, slots, n_slots);
  return 1;
}

int launch_lines(cudaStream_t s, const VmProgram& prog0, const VmProgram& prog1, const double* coord0,
                 const double* coord1, int P, int N, const double* slots, int n_slots, double* T0, int n0, double* T1,
                 int n1) {
  if ((n0 == 0 && n1 == 0) || N == 0) return 0;
  size_t sm = vm_smem(prog0, TB);
  size_t sm1 = vm_smem(prog1, TB);
  if (sm1 > sm) sm = sm1;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_lines), sm, "the per-line part of the expressions")) return 0;
  k_lines<<<blocks_for(static_cast<long long>(N) * P), TB, sm, s>>>(prog0, prog1, coord0, coord1, P, N, slots, n_slots,
                                                                     T0, n0, T1, n1);
  return 1;
}

int launch_cell_moments(cudaStream_t s, const MomentPrograms& mp, MeshDims md, Tables tb, const double* qtab,
                        const double* qsum, int k0, int nk, double stiffness_scale, double* cell, int* error_flag) {
  if (nk == 0) return 0;
  const bool scalar_only = std::getenv("MSGPU_SCALAR_VM") != nullptr;  // A/B switch for measurements
  int r = -1;
  if (!scalar_only) {
    if (md.q == 12) r = dispatch_moments<12, 6, 12>(s, mp, md, tb, k0, nk, stiffness_scale, cell, error_flag);
    else if (md.q == 3) r = dispatch_moments<3, 3, 3>(s, mp, md, tb, k0, nk, stiffness_scale, cell, error_flag);
    else if (md.q == 2) r = dispatch_moments<2, 2, 2>(s, mp, md, tb, k0, nk, stiffness_scale, cell, error_flag);
  }
  if (r >= 0) return r;
  // generic quadrature degree (or a program with very many temporaries): one VM run per point
  const VmProgram& prog = mp.has_map ? *mp.full : *mp.G;
  (void)qsum;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_cell_moments), vm_smem(prog, TB), "the volume integrand")) return 0;
  k_cell_moments<<<blocks_for(static_cast<long long>(nk) * md.ncells), TB, vm_smem(prog, TB), s>>>(
      prog, md, tb, qtab, k0, nk, stiffness_scale, mp.has_map, cell, error_flag);
  return 1;
}

int upload_constant_tables(int q, const double* qtab, const double* qsum, const double* q1d) {
  cudaError_t e = cudaSuccess;
  const size_t bytes = static_cast<size_t>(q) * q * QTAB_STRIDE * sizeof(double), sb = QTAB_STRIDE * sizeof(double);
  const size_t lb = static_cast<size_t>(Q1D_ROWS) * q * sizeof(double);
  if (q == 12) {
    e = cudaMemcpyToSymbol(c_qtab12, qtab, bytes);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_qsum12, qsum, sb);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_q1d12, q1d, lb);
  } else if (q == 3) {
    e = cudaMemcpyToSymbol(c_qtab3, qtab, bytes);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_qsum3, qsum, sb);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_q1d3, q1d, lb);
  } else if (q == 2) {
    e = cudaMemcpyToSymbol(c_qtab2, qtab, bytes);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_qsum2, qsum, sb);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_q1d2, q1d, lb);
  }
  return e == cudaSuccess ? 0 : 1;
}

int launch_face_moments(cudaStream_t s, const VmProgram& prog, int side, MeshDims md, Tables tb, const double* q1d_pt,
                        const double* q1d_wt, int k0, int nk, int has_map, double* face) {
  if (nk == 0) return 0;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_face_moments), vm_smem(prog, TB), "a boundary-face integrand")) return 0;
  k_face_moments<<<blocks_for(static_cast<long long>(nk) * md.nc), TB, vm_smem(prog, TB), s>>>(
      prog, side, md, tb, q1d_pt, q1d_wt, k0, nk, has_map, face);
  return 1;
}

int launch_face_moments_all(cudaStream_t s, const VmProgram* const prog[4], MeshDims md, Tables tb,
                            const double* q1d_pt, const double* q1d_wt, int k0, int nk, int has_map, double* face) {
  if (nk == 0) return 0;
  size_t sm = 0;
  for (int i = 0; i < 4; ++i) sm = vm_smem(*prog[i], TB) > sm ? vm_smem(*prog[i], TB) : sm;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_face_moments4), sm, "a boundary-face integrand")) return 0;
  dim3 grid(blocks_for(static_cast<long long>(nk) * md.nc), 4);
  k_face_moments4<<<grid, TB, sm, s>>>(*prog[0], *prog[1], *prog[2], *prog[3], md, tb, q1d_pt, q1d_wt, k0, nk, has_map,
                                        face);
  return 1;
}

int launch_gather(cudaStream_t s, const GatherArgs& a) {
  if (a.nk == 0) return 0;
  const dim3 grid(static_cast<unsigned>(a.nk), static_cast<unsigned>((a.md.n + TB - 1) / TB));
  if (a.uniform_stiffness != 0)
    k_gather<true><<<grid, TB, 0, s>>>(a);
  else
    k_gather<false><<<grid, TB, 0, s>>>(a);
  return 1;
}

int launch_boundary_values(cudaStream_t s, const VmProgram& prog, MeshDims md, Tables tb, int N, double* bv) {
  if (N == 0) return 0;
  if (!vm_prepare_launch(reinterpret_cast<const void*>(k_boundary_values), vm_smem(prog, TB), "the Dirichlet data")) return 0;
  k_boundary_values<<<blocks_for(static_cast<long long>(N) * 4 * md.nc), TB, vm_smem(prog, TB), s>>>(prog, md, tb, N,
                                                                                                   bv);
  return 1;
}

int launch_rhs_and_dirichlet(cudaStream_t s, const RhsArgs& a) {
  if (a.N == 0) return 0;
  const dim3 blocks(static_cast<unsigned>(a.N), static_cast<unsigned>((a.md.n + TB - 1) / TB));  // (system, node block)
  k_rhs<<<blocks, TB, 0, s>>>(a);
  if (a.problem == PF_ELLIPTIC) {
    k_mask_dirichlet<<<blocks, TB, 0, s>>>(a.md, a.N, a.diag);
    return 2;
  }
  return 1;
}

}  // namespace msgpu
