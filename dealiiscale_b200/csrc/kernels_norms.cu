// The stop test of the fixed-point loop on the device: the macro-aggregated norm of per-grid scalars.
//
// Replaces, for the drivers that keep their macro system on the device,
//   MicroSolver::compute_all_residuals   src/bio-multiscale/micro.cpp:340-365  (per-grid ||v_k - v_k_old||, then the
//                                        L2 norm over the macro mesh of the nodal function with those values)
//   MicroSolver::compute_all_errors      src/elliptic-elliptic/micro.cpp:206-214, src/bio-multiscale/micro.cpp:339-356
//   the nodal norms of MacroSolver::compute_residual (src/bio-multiscale/macro.cpp:230-245)
// all of which are   sqrt( sum_cells ( sqrt( sum_q (sum_a v[dof_a] phi_a(q))^2 JxW ) )^2 )
// (VectorTools::integrate_difference against zero + compute_global_error).  In round 1 the per-grid scalars were
// copied to the host, allgathered from host buffers and summed by host threads: 1 ms of a 12 ms cycle on one GPU,
// 2.5 of 4.9 ms on eight.  Here they never leave the device: k_scatter_sqrt puts sqrt(sum_k) into this rank's block
// of the gather buffer, ncclAllGather runs on that buffer in place, k_nodal_cells integrates one macro cell per thread
// with the host's arithmetic (no FMA contraction, same association, same order of points) and k_sum_in_order adds
// the cells in the host's cell order - one chain of 4096 additions fed from shared memory - so the value is the host's bit for bit and
// every rank takes the same stop decision.  One 8-byte copy comes back.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/msgpu.h"
#include "ctx.h"
#include "fe_tables.h"
#include "kernels.h"

namespace msgpu {
int comm_block(const msgpu_ctx* ctx);
int comm_allgather_macro_buffer(msgpu_ctx* ctx, int fields);
int launch_grid_residual_sums(msgpu_ctx* ctx);  // kernels_extra.cu: d_err[2k] = ||x_k - x_k_old||^2
int launch_grid_error_sums(msgpu_ctx* ctx);     // kernels_extra.cu: d_err[2k], d_err[2k+1] = L2^2, H1-seminorm^2
}  // namespace msgpu
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

// dst[f][k] = sqrt(sums[2 k + f]) (+ sums[2k] under the root for the full H1 norm of the parabolic driver)
__global__ void __launch_bounds__(128) k_scatter_sqrt(int N, int fields, int full_h1, const double* __restrict__ sums,
                                                       double* __restrict__ dst, size_t field_stride) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= N * fields) return;
  const int f = t / N, k = t - f * N;
  double v = sums[2 * k + f];
  if (f == 1 && full_h1) v = __dadd_rn(sums[2 * k], v);
  dst[static_cast<size_t>(f) * field_stride + k] = sqrt(v);
}

// cell_sq[c] = (sqrt(sum_q f_q^2 w_i w_j h h))^2 with f_q = sum_a phi_a(q) v[dof_a]: nodal_l2_norm of host/macro_fem.h,
// operation for operation (products and sums rounded separately, as the host compiles with -ffp-contract=off)
__global__ void __launch_bounds__(128) k_nodal_cells(int n_cells, const int4* __restrict__ cell_dofs,
                                                      const double* __restrict__ values, int q,
                                                      const double* __restrict__ pt, const double* __restrict__ wt,
                                                      double h, double* __restrict__ cell_sq) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= n_cells) return;
  const int4 cd = cell_dofs[c];
  const double u0 = values[cd.x], u1 = values[cd.y], u2 = values[cd.z], u3 = values[cd.w];
  double s = 0.0;
  for (int j = 0; j < q; ++j)
    for (int i = 0; i < q; ++i) {
      const double xi = pt[i], eta = pt[j];
      const double a = __dsub_rn(1.0, xi), b = __dsub_rn(1.0, eta);
      const double v0 = __dmul_rn(a, b), v1 = __dmul_rn(xi, b), v2 = __dmul_rn(a, eta), v3 = __dmul_rn(xi, eta);
      double f = 0.0;
      f = __dadd_rn(f, __dmul_rn(v0, u0));
      f = __dadd_rn(f, __dmul_rn(v1, u1));
      f = __dadd_rn(f, __dmul_rn(v2, u2));
      f = __dadd_rn(f, __dmul_rn(v3, u3));
      const double term = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(f, f), wt[i]), wt[j]), h), h);
      s = __dadd_rn(s, term);
    }
  const double cell = sqrt(s);
  cell_sq[c] = __dmul_rn(cell, cell);
}

// out[f] = sqrt(sum of cell_sq[f][0 .. n_cells) in index order), one CTA per field.  The additions are those of the
// host loop, one dependent chain in index order; the CTA only stages the values in shared memory, 1024 at a time, so
// that the chain waits for the FP64 adder and not for one global load per term (4096 cells: 163 -> ~20 us).
constexpr int SUM_TILE = 1024;
__global__ void __launch_bounds__(256) k_sum_in_order(int n_cells, const double* __restrict__ cell_sq,
                                                       double* __restrict__ out) {
  __shared__ double buf[2][SUM_TILE];
  const int f = blockIdx.x;
  const double* v = cell_sq + static_cast<size_t>(f) * n_cells;
  double total = 0.0;
  int stage = 0;
  for (int i = threadIdx.x; i < SUM_TILE && i < n_cells; i += blockDim.x) buf[0][i] = v[i];
  __syncthreads();
  for (int c0 = 0; c0 < n_cells; c0 += SUM_TILE, stage ^= 1) {
    const int len = n_cells - c0 < SUM_TILE ? n_cells - c0 : SUM_TILE;
    if (threadIdx.x == 0) {
      const double* b = buf[stage];
      int c = 0;
      for (; c + 8 <= len; c += 8) {  // eight loads in flight, one chain of additions
        const double t0 = b[c], t1 = b[c + 1], t2 = b[c + 2], t3 = b[c + 3], t4 = b[c + 4], t5 = b[c + 5], t6 = b[c + 6],
                     t7 = b[c + 7];
        total = __dadd_rn(total, t0);
        total = __dadd_rn(total, t1);
        total = __dadd_rn(total, t2);
        total = __dadd_rn(total, t3);
        total = __dadd_rn(total, t4);
        total = __dadd_rn(total, t5);
        total = __dadd_rn(total, t6);
        total = __dadd_rn(total, t7);
      }
      for (; c < len; ++c) total = __dadd_rn(total, b[c]);
    } else {  // the other threads fetch the next tile meanwhile
      const int n0 = c0 + SUM_TILE;
      for (int i = threadIdx.x - 1; i < SUM_TILE && n0 + i < n_cells; i += blockDim.x - 1) buf[stage ^ 1][i] = v[n0 + i];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[f] = sqrt(total);
}

int norms_of_device_values(msgpu_ctx* ctx, const double* d_values, size_t field_stride, int fields, int q, double* out) {
  NormDevice& Nd = ctx->norms;
  if (q < 1 || q > 16) {
    ctx->err = "nodal norm: quadrature degree 1..16";
    return MSGPU_ERR_INVALID;
  }
  if (Nd.q != q) {
    std::vector<double> pt, wt;
    gauss_legendre_unit(q, pt, wt);
    CK(cudaMemcpyAsync(Nd.d_pt, pt.data(), sizeof(double) * q, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(Nd.d_wt, wt.data(), sizeof(double) * q, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // pt / wt go out of scope
    Nd.q = q;
  }
  for (int f = 0; f < fields; ++f)
    k_nodal_cells<<<(Nd.n_cells + 127) / 128, 128, 0, ctx->stream>>>(Nd.n_cells, Nd.d_cells, d_values + f * field_stride, q,
                                                                     Nd.d_pt, Nd.d_wt, Nd.h,
                                                                     Nd.d_cell_sq + static_cast<size_t>(f) * Nd.n_cells);
  k_sum_in_order<<<fields, 256, 0, ctx->stream>>>(Nd.n_cells, Nd.d_cell_sq, Nd.d_out);
  ctx->launches += fields + 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, Nd.d_out, sizeof(double) * fields, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MSGPU_OK;
}

// per-grid sums in d_err -> sqrt -> this rank's block of the gather buffer -> all ranks' blocks -> norms
int norms_of_grid_sums(msgpu_ctx* ctx, int fields, int full_h1, int q, double* out) {
  NormDevice& Nd = ctx->norms;
  const bool many = ctx->comm_ready && ctx->n_ranks > 1;
  if (many && !ctx->nccl_comm) {
    ctx->err = "nodal norm on the device needs the NCCL communicator (ranks that share a device use the host path)";
    return MSGPU_ERR_UNSUPPORTED;
  }
  const size_t stride = many ? static_cast<size_t>(comm_block(ctx)) * ctx->n_ranks : static_cast<size_t>(Nd.n_values);
  double* buf = many ? ctx->d_macro : Nd.d_values;
  const size_t mine = many ? static_cast<size_t>(ctx->rank) * comm_block(ctx) : 0;
  k_scatter_sqrt<<<(ctx->N * fields + 127) / 128, 128, 0, ctx->stream>>>(ctx->N, fields, full_h1, ctx->d_err, buf + mine,
                                                                         stride);
  ++ctx->launches;
  if (many) {
    const int rc = comm_allgather_macro_buffer(ctx, fields);
    if (rc) return rc;
  }
  return norms_of_device_values(ctx, buf, stride, fields, q, out);
}

}  // namespace

namespace msgpu {
void norms_release(msgpu_ctx* ctx) {
  NormDevice& Nd = ctx->norms;
  void* ptrs[] = {Nd.d_cells, Nd.d_values, Nd.d_cell_sq, Nd.d_out, Nd.d_pt, Nd.d_wt};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  Nd = NormDevice();
}
}  // namespace msgpu

extern "C" {

int msgpu_norm_setup(msgpu_ctx* ctx, const int* cell_dofs, int n_cells, int n_values, double h) {
  if (!ctx || !ctx->setup_done || !cell_dofs || n_cells < 1 || n_values < 1 || !(h > 0.0)) return MSGPU_ERR_INVALID;
  const int ntot = ctx->comm_ready ? ctx->n_total : ctx->N;
  if (n_values != ntot) {
    ctx->err = "nodal norm: one value per grid location (" + std::to_string(n_values) + " != " + std::to_string(ntot) + ")";
    return MSGPU_ERR_INVALID;
  }
  for (int i = 0; i < 4 * n_cells; ++i)
    if (cell_dofs[i] < 0 || cell_dofs[i] >= n_values) {
      ctx->err = "nodal norm: cell_dofs out of range";
      return MSGPU_ERR_INVALID;
    }
  CK(cudaSetDevice(ctx->device));
  norms_release(ctx);
  NormDevice& Nd = ctx->norms;
  Nd.n_cells = n_cells;
  Nd.n_values = n_values;
  Nd.h = h;
  CK(cudaMalloc(reinterpret_cast<void**>(&Nd.d_cells), sizeof(int4) * n_cells));
  CK(cudaMalloc(reinterpret_cast<void**>(&Nd.d_values), sizeof(double) * 2 * n_values));
  CK(cudaMalloc(reinterpret_cast<void**>(&Nd.d_cell_sq), sizeof(double) * 2 * n_cells));
  CK(cudaMalloc(reinterpret_cast<void**>(&Nd.d_out), sizeof(double) * 2));
  CK(cudaMalloc(reinterpret_cast<void**>(&Nd.d_pt), sizeof(double) * 16));
  CK(cudaMalloc(reinterpret_cast<void**>(&Nd.d_wt), sizeof(double) * 16));
  CK(cudaMemcpyAsync(Nd.d_cells, cell_dofs, sizeof(int4) * n_cells, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  Nd.ready = 1;
  return MSGPU_OK;
}

int msgpu_norm_of_values(msgpu_ctx* ctx, const double* values, int q, double* out) {
  if (!ctx || !ctx->norms.ready || !values || !out) return MSGPU_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  NormDevice& Nd = ctx->norms;
  CK(cudaMemcpyAsync(Nd.d_values, values, sizeof(double) * Nd.n_values, cudaMemcpyHostToDevice, ctx->stream));
  return norms_of_device_values(ctx, Nd.d_values, Nd.n_values, 1, q, out);
}

int msgpu_norm_of_residuals(msgpu_ctx* ctx, int q, double* out) {
  if (!ctx || !ctx->norms.ready || !out) return MSGPU_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  const int rc = launch_grid_residual_sums(ctx);
  if (rc) return rc;
  return norms_of_grid_sums(ctx, 1, 0, q, out);
}

int msgpu_norm_of_errors(msgpu_ctx* ctx, int q, double* l2, double* h1) {
  if (!ctx || !ctx->norms.ready || !l2 || !h1) return MSGPU_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  const int rc = launch_grid_error_sums(ctx);
  if (rc) return rc;
  double both[2] = {0.0, 0.0};
  const int rc2 = norms_of_grid_sums(ctx, 2, ctx->desc.problem == MSGPU_PARABOLIC ? 1 : 0, q, both);
  *l2 = both[0];
  *h1 = both[1];
  return rc2;
}

}  // extern "C"
