// Batched FP64 conjugate gradients on sm_100a (kernels K3/K4) and the coupling integrals (K5/K6).
//
// Replaces SolverCG<>::solve(A_k, x_k, b_k, PreconditionIdentity()) called once per micro system
// (reference: src/elliptic-elliptic/micro.cpp:175-182, src/bio-multiscale/micro.cpp:309-314,
// src/elliptic-parabolic/rho_solver.cpp:222-229) and the micro-reading loops of the macro solvers
// (src/elliptic-elliptic/macro.cpp:134-166, src/bio-multiscale/macro.cpp:254-306,
// src/elliptic-parabolic/pi_solver.cpp:178-197).
//
// Design (B200-first): one CTA owns one system at a time and keeps it ON CHIP for the whole
// solve.  A Q1 system on a square mesh is a symmetric 9-point stencil, stored as 5 diagonals
// (offsets 0, +1, +m-1, +m, +m+1 in the y-fastest node ordering).  For m <= 67 the 5 diagonals
// plus the search direction d fit in the 227 KB of shared memory of one SM (206 KB at m = 65);
// x, g and the thread's own d entries live in registers.  The matrix is therefore read from
// HBM exactly once per solve instead of once per CG iteration, there are no column indices at
// all, and every shared-memory access is unit-stride (row r = thread + i*blockDim).
// CTAs pull systems from an atomic work counter, so different iteration counts balance out.
//
// The recurrence is deal.II 9.1's SolverCG: g = A x - b; d = -g; loop { h = A d;
// alpha = gh/(d.h); x += alpha d; g += alpha h; res = |g|; stop if res <= tol;
// beta = res^2/gh; d = beta d - g }.  Dot products use a fixed shuffle tree + fixed order over
// warps, so results are bit-reproducible from run to run and independent of the SM.
#include <cooperative_groups.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "cg_common.cuh"
#include "checked.cuh"
#include "kernels.h"

namespace msgpu {

namespace {

using cgc::l2_prefetch_bulk;
using cgc::strip_dot;
using cgc::warp_sum;

// Sum over the CTA, result in every thread.  One barrier; `red` must not be reused before the
// next barrier that follows this call (callers alternate between buffers).
template <int BT>
__device__ __forceinline__ double block_sum(double v, double* red) {
  constexpr int NW = BT / 32;
  v = warp_sum(v);
  if (NW == 1) return v;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = (lane < NW) ? red[lane] : 0.0;
  return warp_sum(t);
}

// Same contract, shorter dependency chain for small warp counts: every thread reads all NW warp
// sums (broadcast loads) and adds them in one fixed binary tree instead of a second shuffle tree.
template <int BT>
__device__ __forceinline__ double block_sum_tree(double v, double* red) {
  constexpr int NW = BT / 32;
  static_assert(NW >= 2 && NW <= 16 && NW % 2 == 0, "warp count");
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t[16];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    if (i < NW) {
      const double2 p = *reinterpret_cast<const double2*>(red + i);
      t[i] = p.x;
      t[i + 1] = p.y;
    } else {
      t[i] = 0.0;
      t[i + 1] = 0.0;
    }
  }
#pragma unroll
  for (int w2 = 1; w2 < 16; w2 *= 2)
#pragma unroll
    for (int i = 0; i + w2 < 16; i += 2 * w2)
      if (i + w2 < NW) t[i] += t[i + w2];
  return t[0];
}

#ifdef MSGPU_AB_VARIANTS  // superseded generation, kept for A/B measurements (scripts/cg_variants.py): -DMSGPU_AB_VARIANTS
// -------------------------------------------------------------------------------------------
// Shared-memory resident CG.  BT threads, RPT rows per thread (BT*RPT >= n).
template <int BT, int RPT, bool JACOBI>
__global__ void __launch_bounds__(BT, 1) k_cg_resident(CgArgs a) {
  if (a.gate != nullptr && ((*a.gate == 0) != (a.gate_run_if_zero != 0))) return;  // the other CG kernel does the work
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_k;
  const int n = a.md.n, m = a.md.m, ld = a.md.ld;
  const int pad = (m + 2) & ~1;          // even, >= m+1: zero guard in front of every diagonal
  const int ms = pad + ld;               // stride between diagonals in shared memory
  double* const mat = sm;                // [NDIAG][ms]
  double* const dv = sm + NDIAG * ms + pad;   // dv[-pad .. n + pad)
  double* const red = sm + NDIAG * ms + (n + 2 * pad);   // 3 x 32
  const int tid = threadIdx.x;

  // guards: zero once, never written again
  for (int i = tid; i < NDIAG * ms; i += BT)
    if ((i % ms) < pad || (i % ms) >= pad + n) mat[i] = 0.0;
  for (int i = tid; i < pad; i += BT) {
    dv[-pad + i] = 0.0;
    dv[n + i] = 0.0;
  }
  const double* const M0 = mat + pad;
  const double* const M1 = M0 + ms;
  const double* const M2 = M1 + ms;
  const double* const M3 = M2 + ms;
  const double* const M4 = M3 + ms;
  const int o2 = m - 1, o3 = m, o4 = m + 1;

  auto spmv_row = [&](int r) -> double {
    double s = M0[r] * dv[r];
    s += M1[r] * dv[r + 1];
    s += M1[r - 1] * dv[r - 1];
    s += M2[r] * dv[r + o2];
    s += M2[r - o2] * dv[r - o2];
    s += M3[r] * dv[r + o3];
    s += M3[r - o3] * dv[r - o3];
    s += M4[r] * dv[r + o4];
    s += M4[r - o4] * dv[r - o4];
    return s;
  };

  for (;;) {
    __syncthreads();  // everybody is done with the previous system (s_k, mat, dv)
    if (tid == 0) s_k = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int k = s_k;
    if (k >= a.N) break;

    // ---- stage the matrix: HBM -> shared memory, once per solve
    const double* dg = a.diag + static_cast<size_t>(k) * a.diag_stride;
    for (int d = 0; d < NDIAG; ++d) {
      const double* src = dg + static_cast<size_t>(d) * ld;
      double* dst = mat + d * ms + pad;
      for (int r = tid; r < n; r += BT) dst[r] = __ldg(src + r);
    }
    double xx[RPT], gg[RPT], dl[RPT];
    const double* bk = a.b + static_cast<size_t>(k) * n;
    double* xk = a.x + static_cast<size_t>(k) * n;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = tid + i * BT;
      xx[i] = (r < n) ? xk[r] : 0.0;
      if (r < n) dv[r] = xx[i];
    }
    __syncthreads();
    // ---- g = A x - b
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = tid + i * BT;
      gg[i] = (r < n) ? (spmv_row(r) - bk[r]) : 0.0;
      part += gg[i] * gg[i];
    }
    double res = sqrt(block_sum<BT>(part, red));
    int it = 0;
    bool ok = res <= a.tol;
    double gh = 0.0;
    if (!ok) {
      __syncthreads();  // all rows of A x are done before d is overwritten
      if (JACOBI) {
        part = 0.0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = tid + i * BT;
          const double z = (r < n) ? gg[i] / M0[r] : 0.0;
          dl[i] = -z;
          part += gg[i] * z;
          if (r < n) dv[r] = dl[i];
        }
        gh = block_sum<BT>(part, red + 32);
      } else {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = tid + i * BT;
          dl[i] = -gg[i];
          if (r < n) dv[r] = dl[i];
        }
        gh = res * res;
      }
      for (;;) {
        ++it;
        __syncthreads();  // d complete
        double hh[RPT];
        part = 0.0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = tid + i * BT;
          hh[i] = (r < n) ? spmv_row(r) : 0.0;
          part += dl[i] * hh[i];
        }
        const double alpha = gh / block_sum<BT>(part, red);
        part = 0.0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          xx[i] += alpha * dl[i];
          gg[i] += alpha * hh[i];
          part += gg[i] * gg[i];
        }
        res = sqrt(block_sum<BT>(part, red + 32));
        if (res <= a.tol) {
          ok = true;
          break;
        }
        if (it >= a.max_it || res != res) break;
        double beta = gh;
        if (JACOBI) {
          part = 0.0;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = tid + i * BT;
            hh[i] = (r < n) ? gg[i] / M0[r] : 0.0;
            part += gg[i] * hh[i];
          }
          gh = block_sum<BT>(part, red + 64);
          beta = gh / beta;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = tid + i * BT;
            dl[i] = beta * dl[i] - hh[i];
            if (r < n) dv[r] = dl[i];
          }
        } else {
          gh = res * res;
          beta = gh / beta;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = tid + i * BT;
            dl[i] = beta * dl[i] - gg[i];
            if (r < n) dv[r] = dl[i];
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = tid + i * BT;
      if (r < n) xk[r] = xx[i];
    }
    if (tid == 0) {
      if (a.iters) a.iters[k] = it;
      if (a.res) a.res[k] = res;
      if (!ok) atomicExch(a.fail_flag, 1);
    }
  }
}


#endif  // MSGPU_AB_VARIANTS

// -------------------------------------------------------------------------------------------
// Shared-memory resident CG, strip ownership: thread t owns the K consecutive rows t*K .. t*K+K-1
// (consecutive nodes of one mesh column in the y-fastest ordering).  Compared with the strided
// variant above this cuts the shared-memory traffic per row from 17 to ~10.8 doubles:
//   * the +-1 neighbours inside the strip come from registers,
//   * the off-diagonal M1[r] is loaded once for rows r and r+1,
//   * the left / right column values slide through a 3-register window.
// The stride between lanes is K doubles; K is odd, so 64-bit accesses stay bank-conflict free.
// PRE: 0 = PreconditionIdentity, 1 = Jacobi, 2 = SSOR.  The SSOR sweeps run in the four-colour ordering of the
// lattice (colour = parity of (ix, iy)): in a 9-point stencil no two nodes of one colour are coupled, so a sweep is
// four fully parallel stages instead of n sequential updates — M = (D/w + L)(D/w)^-1 (D/w + U)/(2 - w) with L / U
// the couplings to lower / higher colours, symmetric positive definite like the lexicographic SSOR it replaces.
// Vertical neighbours of a node of colour c have colour c^2, horizontal ones c^1, diagonal ones c^3.
template <int BT, int K, int PRE>
__global__ void __launch_bounds__(BT, 1) k_cg_strip(CgArgs a) {
  if (a.gate != nullptr && ((*a.gate == 0) != (a.gate_run_if_zero != 0))) return;  // the other CG kernel does the work
  constexpr bool JACOBI = PRE == 1;
  constexpr bool SSOR = PRE == 2;
  static_assert(K % 2 == 1, "odd strip length keeps the shared-memory accesses conflict free");
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_k;
  const int n = a.md.n, m = a.md.m, ld = a.md.ld;
  const int pad = (m + 2) & ~1;
  const int ms = pad + ld;
  double* const mat = sm;
  double* const dv = sm + NDIAG * ms + pad;
  double* const red = sm + NDIAG * ms + (n + 2 * pad);
  const int tid = threadIdx.x;
  const int r0 = tid * K;
  unsigned long long colours = 0;  // two bits per row of the strip
  if (SSOR) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = r0 + j, ix = r / m, iy = r - ix * m;
      colours |= static_cast<unsigned long long>((ix & 1) | ((iy & 1) << 1)) << (2 * j);
    }
  }

  for (int i = tid; i < NDIAG * ms; i += BT)
    if ((i % ms) < pad || (i % ms) >= pad + n) mat[i] = 0.0;
  for (int i = tid; i < pad; i += BT) {
    dv[-pad + i] = 0.0;
    dv[n + i] = 0.0;
  }
  const double* const M0 = mat + pad;
  const double* const M1 = M0 + ms;
  const double* const M2 = M1 + ms;
  const double* const M3 = M2 + ms;
  const double* const M4 = M3 + ms;

  // h = A d for the K rows of this thread; dl = own entries of d (registers), dv = all of d (shared)
  auto spmv_strip = [&](const double (&dl)[K], double (&hh)[K]) {
    if (r0 >= n) {
#pragma unroll
      for (int j = 0; j < K; ++j) hh[j] = 0.0;
      return;
    }
    // sliding windows over the right (r+m-1, r+m, r+m+1) and left (r-m-1, r-m, r-m+1) columns
    double ra = dv[r0 + m - 1], rb = dv[r0 + m];
    double la = dv[r0 - m - 1], lb = dv[r0 - m];
    double below = dv[r0 - 1];   // d[r-1] of the first row
    double m1_prev = M1[r0 - 1];  // coupling (r-1, r)
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = r0 + j;
      if (r < n) {
        const double rc = dv[r + m + 1], lc = dv[r - m + 1];
        const double m1 = M1[r];
        const double above = (j + 1 < K) ? dl[j + 1] : dv[r + 1];
        double s = M0[r] * dl[j];
        s += m1 * above;
        s += m1_prev * below;
        s += M2[r] * ra;
        s += M2[r - (m - 1)] * lc;
        s += M3[r] * rb;
        s += M3[r - m] * lb;
        s += M4[r] * rc;
        s += M4[r - (m + 1)] * la;
        hh[j] = s;
        ra = rb; rb = rc;
        la = lb; lb = lc;
        below = dl[j];
        m1_prev = m1;
      } else {
        hh[j] = 0.0;
      }
    }
  };

  // z = M^-1 g for the rows of this thread (SSOR); dv is the work space of the sweeps and holds garbage
  // afterwards — every caller rewrites it with the new search direction behind the closing barrier
  auto ssor_apply = [&](const double (&g_)[K], double (&z_)[K]) {
    const double w = a.omega;
    auto coupled = [&](int r, int c, bool lower) -> double {
      const bool useV = lower ? ((c ^ 2) < c) : ((c ^ 2) > c);
      const bool useH = lower ? ((c ^ 1) < c) : ((c ^ 1) > c);
      const bool useD = lower ? ((c ^ 3) < c) : ((c ^ 3) > c);
      double s = 0.0;
      if (useV) s += M1[r] * dv[r + 1] + M1[r - 1] * dv[r - 1];
      if (useH) s += M3[r] * dv[r + m] + M3[r - m] * dv[r - m];
      if (useD)
        s += M2[r] * dv[r + m - 1] + M2[r - (m - 1)] * dv[r - (m - 1)] + M4[r] * dv[r + m + 1] +
             M4[r - (m + 1)] * dv[r - (m + 1)];
      return s;
    };
    for (int c = 0; c < 4; ++c) {  // forward: (D/w + L) y = g
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int r = r0 + j;
        if (r < n && static_cast<int>((colours >> (2 * j)) & 3u) == c) {
          const double y = (g_[j] - w * coupled(r, c, true)) / M0[r];
          z_[j] = y;
          dv[r] = y;
        }
      }
      __syncthreads();
    }
    for (int c = 2; c >= 0; --c) {  // backward: (D/w + U) z = (D/w) y; colour 3 has nothing above it
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int r = r0 + j;
        if (r < n && static_cast<int>((colours >> (2 * j)) & 3u) == c) {
          const double z = z_[j] - w * coupled(r, c, false) / M0[r];
          z_[j] = z;
          dv[r] = z;
        }
      }
      __syncthreads();
    }
    const double scale = w * (2.0 - w);
#pragma unroll
    for (int j = 0; j < K; ++j) z_[j] = (r0 + j < n) ? z_[j] * scale : 0.0;
  };

  bool have_matrix = false;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_k = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int k = s_k;
    if (k >= a.N) break;

    // one matrix shared by all systems (diag_stride == 0: the parabolic driver, many right-hand sides): it is
    // staged once per CTA and stays resident while the CTA walks through its systems
    if (a.diag_stride != 0 || !have_matrix) {
      const double* dg = a.diag + static_cast<size_t>(k) * a.diag_stride;
      for (int d = 0; d < NDIAG; ++d) {
        const double* src = dg + static_cast<size_t>(d) * ld;
        double* dst = mat + d * ms + pad;
        for (int r = tid; r < n; r += BT) dst[r] = __ldg(src + r);
      }
      have_matrix = true;
    }
    double xx[K], gg[K], dl[K], hh[K];
    const double* bk = a.b + static_cast<size_t>(k) * n;
    double* xk = a.x + static_cast<size_t>(k) * n;
    // coalesced staging of the start vector through dv, then every thread picks up its strip
    for (int r = tid; r < n; r += BT) dv[r] = xk[r];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = r0 + j;
      xx[j] = (r < n) ? dv[r] : 0.0;
      dl[j] = xx[j];
    }
    // ---- g = A x - b
    spmv_strip(dl, hh);
    double part = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = r0 + j;
      gg[j] = (r < n) ? (hh[j] - bk[r]) : 0.0;
      part += gg[j] * gg[j];
    }
    double res = sqrt(block_sum<BT>(part, red));
    int it = 0;
    bool ok = res <= a.tol;
    double gh = 0.0;
    if (!ok) {
      __syncthreads();
      if (SSOR) {
        ssor_apply(gg, hh);
        part = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int r = r0 + j;
          dl[j] = -hh[j];
          part += gg[j] * hh[j];
          if (r < n) dv[r] = dl[j];
        }
        gh = block_sum<BT>(part, red + 32);
      } else if (JACOBI) {
        part = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int r = r0 + j;
          const double z = (r < n) ? gg[j] / M0[r] : 0.0;
          dl[j] = -z;
          part += gg[j] * z;
          if (r < n) dv[r] = dl[j];
        }
        gh = block_sum<BT>(part, red + 32);
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int r = r0 + j;
          dl[j] = -gg[j];
          if (r < n) dv[r] = dl[j];
        }
        gh = res * res;
      }
      for (;;) {
        ++it;
        __syncthreads();
        spmv_strip(dl, hh);
        part = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) part += dl[j] * hh[j];
        const double alpha = gh / block_sum<BT>(part, red);
        part = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          xx[j] += alpha * dl[j];
          gg[j] += alpha * hh[j];
          part += gg[j] * gg[j];
        }
        res = sqrt(block_sum<BT>(part, red + 32));
        if (res <= a.tol) {
          ok = true;
          break;
        }
        if (it >= a.max_it || res != res) break;
        double beta = gh;
        if (JACOBI || SSOR) {
          part = 0.0;
          if (SSOR) {
            ssor_apply(gg, hh);
#pragma unroll
            for (int j = 0; j < K; ++j) part += gg[j] * hh[j];
          } else {
#pragma unroll
            for (int j = 0; j < K; ++j) {
              const int r = r0 + j;
              hh[j] = (r < n) ? gg[j] / M0[r] : 0.0;
              part += gg[j] * hh[j];
            }
          }
          gh = block_sum<BT>(part, red + 64);
          beta = gh / beta;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const int r = r0 + j;
            dl[j] = beta * dl[j] - hh[j];
            if (r < n) dv[r] = dl[j];
          }
        } else {
          gh = res * res;
          beta = gh / beta;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const int r = r0 + j;
            dl[j] = beta * dl[j] - gg[j];
            if (r < n) dv[r] = dl[j];
          }
        }
      }
    }
    // coalesced write-back through shared memory
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = r0 + j;
      if (r < n) dv[r] = xx[j];
    }
    __syncthreads();
    for (int r = tid; r < n; r += BT) xk[r] = dv[r];
    if (tid == 0) {
      if (a.iters) a.iters[k] = it;
      if (a.res) a.res[k] = res;
      if (!ok) atomicExch(a.fail_flag, 1);
    }
  }
}


// -------------------------------------------------------------------------------------------
// Strip CG with the matrix in TENSOR MEMORY (TMEM).
//
// The strip kernel above is limited by the shared-memory pipe, and 8 of its ~11 loads per row
// are matrix entries.  Seven of them are needed by the thread that owns row r and by nobody else:
// M0[r], M2[r], M3[r], M4[r] and the transposed M2[r-(m-1)], M3[r-m], M4[r-(m+1)] — a lane-private
// pattern, which is exactly what TMEM offers: 128 lanes x 512 32-bit columns per SM behind its own
// datapath (tcgen05.ld / tcgen05.st), unused otherwise in this FP64 kernel.  Warp w may touch lanes
// 32*(w%4)..+31; the BT/128 warps of a lane quadrant split the 512 columns, so a thread has
// 64 (BT = 512) or 128 (BT = 256) doubles of TMEM.  Row j of the strip keeps its 7 entries in
// columns 14j..14j+13 and fetches them with ONE tcgen05.ld.32x32b.x16 per row per iteration (the
// two surplus columns belong to the next row / are never used), software-pipelined one row ahead.
// Shared memory then only serves M1 and the search direction (~3.7 instead of ~11.8 loads per
// row); M2..M4 are staged there once per system for the transposed pick-up.
// Rows past n are handled without branches: their matrix entries are exact zeros and dead threads
// work on the zero guard behind d, so the compiler is free to interleave the rows of a strip.
// Every lane of every warp executes the (sync.aligned) tcgen05 instructions.
struct TmemRow {
  uint32_t r[16];
};
__device__ __forceinline__ void tmem_store_row(uint32_t taddr, const double (&v)[7]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};\n"
      "tcgen05.wait::st.sync.aligned;\n" ::"r"(taddr),
      "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
      "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
      "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
      "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(0), "r"(0));
}
// Asynchronous TMEM load of one row (16 columns) and the matching wait.  TMEM is not aliased by any
// pointer, so neither needs a "memory" clobber (which would also fence the shared-memory loads of
// neighbouring rows); the wait takes the destination registers as in/out operands so that no use of
// them can be scheduled before it.
__device__ __forceinline__ void tmem_issue_row(uint32_t taddr, TmemRow& q) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];\n"
      : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]), "=r"(q.r[4]), "=r"(q.r[5]), "=r"(q.r[6]), "=r"(q.r[7]),
        "=r"(q.r[8]), "=r"(q.r[9]), "=r"(q.r[10]), "=r"(q.r[11]), "=r"(q.r[12]), "=r"(q.r[13]), "=r"(q.r[14]),
        "=r"(q.r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_row(TmemRow& q) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(q.r[0]), "+r"(q.r[1]), "+r"(q.r[2]), "+r"(q.r[3]), "+r"(q.r[4]), "+r"(q.r[5]), "+r"(q.r[6]),
                 "+r"(q.r[7]), "+r"(q.r[8]), "+r"(q.r[9]), "+r"(q.r[10]), "+r"(q.r[11]), "+r"(q.r[12]), "+r"(q.r[13]));
}
__device__ __forceinline__ double tmem_get(const TmemRow& q, int i) { return __hiloint2double(q.r[2 * i + 1], q.r[2 * i]); }

__host__ __device__ constexpr int tmem_pad(int m, int K) { return (m + 3 + K) & ~1; }

// TMEM columns a CTA of BT threads with K rows per thread allocates: the smallest power of two (>= 32) that holds
// 14 K + 2 columns for each of the BT/128 warps of a lane quadrant.  Shapes that need less than all 512 columns
// leave room for further CTAs on the same SM — several small systems in flight per SM, each hiding the
// reduction latency of the others.
__host__ __device__ constexpr int tmem_alloc_columns(int BT, int K) {
  int need = (BT / 128) * (14 * K + 2), c = 32;
  while (c < need) c *= 2;
  return c;
}

template <int BT, int K, bool JACOBI>
__global__ void __launch_bounds__(BT, 512 / tmem_alloc_columns(BT, K)) k_cg_strip_tmem(CgArgs a) {
  if (a.gate != nullptr && ((*a.gate == 0) != (a.gate_run_if_zero != 0))) return;  // the other CG kernel does the work
  constexpr int ALLOC = tmem_alloc_columns(BT, K);
  constexpr int CW = (ALLOC / (BT / 128)) & ~1;  // TMEM columns per warp
  constexpr int NACC = K >= 12 ? 4 : 1;          // K = 9 stays bit-identical to k_cg_strip
  constexpr bool FAST = MSGPU_FAST_SCALARS;
  static_assert(BT % 128 == 0 && ALLOC <= 512 && 14 * K + 2 <= CW, "strip must fit the TMEM columns of a thread");
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_k, s_next;
  __shared__ uint32_t s_tmem;
  const int n = a.md.n, m = a.md.m, ld = a.md.ld;
  const int pad = tmem_pad(m, K);  // >= m + 1 + K: dead rows n..n+K-1 stay inside the guards
  const int ms = pad + ld;
  // shared memory: M1..M4 (4 diagonals) + d with halos + reduction scratch
  // (checked builds, checked.cuh: every access below is compared with the window it may touch)
  const CheckedPtr<double> mat = checked(sm, sm, sm + 4 * ms);
  const CheckedPtr<double> dv = checked(sm + 4 * ms + pad, sm + 4 * ms, sm + 4 * ms + n + 2 * pad);
  double* const red = sm + ((4 * ms + n + 2 * pad + 1) & ~1);  // 16-byte aligned
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r0 = tid * K;
  const int rc0 = r0 < n ? r0 : n;                      // dead threads work on the zero guard
  const int nlive = r0 < n ? (n - r0 < K ? n - r0 : K) : 0;  // rows of this strip inside the system

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
        static_cast<uint32_t>(__cvta_generic_to_shared(&s_tmem))), "n"(ALLOC));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  for (int i = tid; i < 4 * ms; i += BT)
    if ((i % ms) < pad || (i % ms) >= pad + n) mat[i] = 0.0;
  for (int i = tid; i < pad; i += BT) {
    dv[-pad + i] = 0.0;
    dv[n + i] = 0.0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // lane quadrant of this warp in bits 31:16, column block of this warp in bits 15:0
  const uint32_t tbase =
      s_tmem + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + static_cast<uint32_t>(warp >> 2) * static_cast<uint32_t>(CW);

  const CheckedPtr<const double> M1 = mat + pad;
  const CheckedPtr<const double> M2 = M1 + ms;
  const CheckedPtr<const double> M3 = M2 + ms;
  const CheckedPtr<const double> M4 = M3 + ms;
  // tensor memory: rows j = 0 .. K-1 of this thread at columns tbase + 14 j .. + 15, inside the CW columns of the warp
  [[maybe_unused]] const uint32_t tm_lo = tbase & 0xffffu, tm_hi = tm_lo + CW;
  MSGPU_CHECK((warp >> 2) * CW + CW <= ALLOC);

  // h = A d for the K rows of this thread; dl = own entries of d (registers), dv = all of d (shared)
  auto spmv_strip = [&](const double (&dl)[K], double (&hh)[K]) {
    // sliding windows over the right (r+m-1, r+m, r+m+1) and left (r-m-1, r-m, r-m+1) columns
    double ra = dv[rc0 + m - 1], rb = dv[rc0 + m];
    double la = dv[rc0 - m - 1], lb = dv[rc0 - m];
    double below = dv[rc0 - 1];
    double m1_prev = M1[rc0 - 1];
    TmemRow q[2];  // double buffer; the loop is fully unrolled, so q[j & 1] are plain registers
    MSGPU_CHECK_TMEM(tbase, 16u, tm_lo, tm_hi);
    tmem_issue_row(tbase, q[0]);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = rc0 + j;
      TmemRow& cur = q[j & 1];
      tmem_wait_row(cur);
      if (j + 1 < K) {
        MSGPU_CHECK_TMEM(tbase + 14u * (j + 1), 16u, tm_lo, tm_hi);
        tmem_issue_row(tbase + 14u * (j + 1), q[(j + 1) & 1]);  // one row ahead
      }
      const double rc = dv[r + m + 1], lc = dv[r - m + 1];
      const double m1 = M1[r];
      const double above = (j + 1 < K) ? dl[j + 1] : dv[r + 1];
      if (NACC == 1) {
        double s = tmem_get(cur, 0) * dl[j];
        s += m1 * above;
        s += m1_prev * below;
        s += tmem_get(cur, 1) * ra;
        s += tmem_get(cur, 4) * lc;
        s += tmem_get(cur, 2) * rb;
        s += tmem_get(cur, 5) * lb;
        s += tmem_get(cur, 3) * rc;
        s += tmem_get(cur, 6) * la;
        hh[j] = s;
      } else {
        // two shorter dependency chains per row (same terms, fixed order)
        double s = tmem_get(cur, 0) * dl[j];
        double t = tmem_get(cur, 1) * ra;
        s += m1 * above;
        t += tmem_get(cur, 4) * lc;
        s += m1_prev * below;
        t += tmem_get(cur, 2) * rb;
        s += tmem_get(cur, 5) * lb;
        t += tmem_get(cur, 3) * rc;
        s += tmem_get(cur, 6) * la;
        hh[j] = s + t;
      }
      ra = rb; rb = rc;
      la = lb; lb = lc;
      below = dl[j];
      m1_prev = m1;
    }
  };

  // The queue is read one system ahead: while system k is being solved, the five diagonals, the right-hand
  // side and the start vector of the NEXT system are pulled into L2 by one bulk prefetch each
  // (cp.async.bulk.prefetch.L2, issued by one thread, no completion to wait for), so that the staging loops
  // below find them there instead of in HBM.
  if (tid == 0) s_next = atomicAdd(a.work_counter, 1);
  bool have_matrix = false;
  // Jacobi: reciprocal diagonal in the staging area of M2 (free once the matrix sits in tensor memory): no extra
  // registers in the CG loop, a multiplication instead of a division per row
  const CheckedPtr<double> rdiag = mat + (ms + pad);
  double d0[JACOBI ? K : 1];  // transient: the diagonal of this strip between packing and the barrier
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_k = s_next;
      if (s_next < a.N) {
        const int nx = atomicAdd(a.work_counter, 1);
        s_next = nx;
        if (nx < a.N) {
          const uint32_t row_bytes = static_cast<uint32_t>(n) * 8u & ~15u;
          const double* ng = a.diag + static_cast<size_t>(nx) * a.diag_stride;
          if (a.diag_stride != 0)
            for (int d = 0; d < NDIAG; ++d) l2_prefetch_bulk(ng + static_cast<size_t>(d) * ld, row_bytes);
          l2_prefetch_bulk(a.b + static_cast<size_t>(nx) * n, row_bytes);
          l2_prefetch_bulk(a.x + static_cast<size_t>(nx) * n, row_bytes);
        }
      }
    }
    __syncthreads();
    const int k = s_k;
    if (k >= a.N) break;

    // One matrix shared by all systems (diag_stride == 0: the parabolic driver, many right-hand sides): staged
    // and packed into tensor memory once per CTA, resident while the CTA walks through its systems.
    const bool load_matrix = a.diag_stride != 0 || !have_matrix;
    have_matrix = true;
    const double* dg = a.diag + static_cast<size_t>(k) * a.diag_stride;
    // stage M1..M4 in shared memory; M0 goes through the d buffer on its way to TMEM
    if (load_matrix) {
      for (int d = 1; d < NDIAG; ++d) {
        const double* src = dg + static_cast<size_t>(d) * ld;
        const CheckedPtr<double> dst = mat + ((d - 1) * ms + pad);
        for (int r = tid; r < n; r += BT) dst[r] = __ldg(src + r);
      }
      for (int r = tid; r < n; r += BT) dv[r] = __ldg(dg + r);
    }
    __syncthreads();
    if (load_matrix)
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = rc0 + j;
      const bool in = j < nlive;
      double v[7];
      v[0] = in ? dv[r] : 0.0;
      v[1] = in ? M2[r] : 0.0;
      v[2] = in ? M3[r] : 0.0;
      v[3] = in ? M4[r] : 0.0;
      v[4] = in ? M2[r - (m - 1)] : 0.0;
      v[5] = in ? M3[r - m] : 0.0;
      v[6] = in ? M4[r - (m + 1)] : 0.0;
      if (JACOBI) d0[j] = in ? v[0] : 1.0;
      MSGPU_CHECK_TMEM(tbase + 14u * j, 16u, tm_lo, tm_hi);
      tmem_store_row(tbase + 14u * j, v);  // ascending j: the two surplus columns are rewritten by row j+1
    }
    __syncthreads();  // everybody has picked up M0 before d is overwritten
    if (JACOBI && load_matrix) {
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (j < nlive) rdiag[rc0 + j] = 1.0 / d0[j];  // rows past n keep the zero of the guard: z = 0 there
    }

    double xx[K], gg[K], dl[K], hh[K];
    const double* bk = a.b + static_cast<size_t>(k) * n;
    double* xk = a.x + static_cast<size_t>(k) * n;
    for (int r = tid; r < n; r += BT) dv[r] = xk[r];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) {
      xx[j] = (j < nlive) ? dv[r0 + j] : 0.0;
      dl[j] = xx[j];
    }
    spmv_strip(dl, hh);
    double part = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      gg[j] = (j < nlive) ? (hh[j] - bk[r0 + j]) : 0.0;
      part += gg[j] * gg[j];
    }
    double res = sqrt(block_sum_tree<BT>(part, red));
    int it = 0;
    bool ok = res <= a.tol;
    double gh = 0.0;
    if (!ok) {
      __syncthreads();
      if (JACOBI) {
        part = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const double z = gg[j] * rdiag[rc0 + j];
          dl[j] = -z;
          part += gg[j] * z;
          if (j < nlive) dv[r0 + j] = dl[j];
        }
        gh = block_sum_tree<BT>(part, red + 32);
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          dl[j] = -gg[j];
          if (j < nlive) dv[r0 + j] = dl[j];
        }
        gh = res * res;
      }
      double rgh = FAST ? 1.0 / gh : 0.0;
      for (;;) {
        ++it;
        __syncthreads();
        spmv_strip(dl, hh);
        part = strip_dot<K, NACC>([&](int j) { return dl[j] * hh[j]; });
        const double alpha = gh / block_sum_tree<BT>(part, red);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          xx[j] += alpha * dl[j];
          gg[j] += alpha * hh[j];
        }
        part = strip_dot<K, NACC>([&](int j) { return gg[j] * gg[j]; });
        const double s2 = block_sum_tree<BT>(part, red + 32);
        double gh_new;
        if (FAST) {
          res = s2;
          gh_new = s2;
          if (s2 <= a.tol2) {
            ok = true;
            break;
          }
        } else {
          res = sqrt(s2);
          gh_new = res * res;
          if (res <= a.tol) {
            ok = true;
            break;
          }
        }
        if (it >= a.max_it || s2 != s2) break;
        double beta = gh;
        if (JACOBI) {
          part = 0.0;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            hh[j] = gg[j] * rdiag[rc0 + j];
            part += gg[j] * hh[j];
          }
          gh = block_sum_tree<BT>(part, red + 64);
          beta = gh / beta;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            dl[j] = beta * dl[j] - hh[j];
            if (j < nlive) dv[r0 + j] = dl[j];
          }
        } else {
          beta = FAST ? gh_new * rgh : gh_new / beta;
          gh = gh_new;
          if (FAST) rgh = 1.0 / gh;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            dl[j] = beta * dl[j] - gg[j];
            if (j < nlive) dv[r0 + j] = dl[j];
          }
        }
      }
    }
    if (FAST && it > 0) res = sqrt(res);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (j < nlive) dv[r0 + j] = xx[j];
    __syncthreads();
    for (int r = tid; r < n; r += BT) xk[r] = dv[r];
    if (tid == 0) {
      if (a.iters) a.iters[k] = it;
      if (a.res) a.res[k] = res;
      if (!ok) atomicExch(a.fail_flag, 1);
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem), "n"(ALLOC));
}

// -------------------------------------------------------------------------------------------
// Cluster CG for systems that do not fit one SM (m > 67, e.g. the 128 x 128-cell micro meshes).
//
// A thread-block cluster of CL CTAs owns one system; CTA c keeps rows [c*RC, (c+1)*RC) exactly as
// k_cg_strip_tmem keeps a whole system: seven matrix entries per row in its SM's tensor memory, M1
// and its window of the search direction in shared memory, x / g / d / h strips in registers.
// What crosses CTAs:
//   * the halo of d: after every update a CTA stores its first / last `pad` entries straight into
//     the neighbours' shared memory (distributed shared memory, plain stores through
//     cluster.map_shared_rank);
//   * the two dot products per step: every CTA reduces its rows, stores the partial into slot
//     [rank] of every CTA's `cred` array, and after the cluster barrier all CTAs add the CL partials
//     in rank order — every CTA sees bit-identical scalars, so the stop decision is uniform.
// Three cluster barriers per CG step (halo, two reductions) replace the three CTA barriers.
namespace cg = cooperative_groups;

template <int K, bool JACOBI>
__global__ void __launch_bounds__(256, 1) k_cg_cluster(CgArgs a) {
  if (a.gate != nullptr && ((*a.gate == 0) != (a.gate_run_if_zero != 0))) return;  // the other CG kernel does the work
  constexpr int BT = 256, CW = 256, NACC = 4, RC = BT * K;
  static_assert(14 * K + 2 <= CW, "strip must fit the TMEM columns of a thread");
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = static_cast<int>(cluster.num_blocks());
  const int crank = static_cast<int>(cluster.block_rank());
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_k;
  __shared__ uint32_t s_tmem;
  const int n = a.md.n, m = a.md.m, ld = a.md.ld;
  const int pad = tmem_pad(m, K);
  const int win = RC + 2 * pad;            // window of this CTA: local index i in [-pad, RC + pad)
  const int lo = crank * RC;               // first global row of this CTA
  const int nloc = n - lo < RC ? (n - lo > 0 ? n - lo : 0) : RC;
  // (checked builds, checked.cuh: every access, local or in a neighbour's shared memory, is compared with its window)
  const CheckedPtr<double> m1 = checked(sm + pad, sm, sm + win);                   // M1[lo + i]
  const CheckedPtr<double> dv = checked(sm + win + pad, sm + win, sm + 2 * win);  // d[lo + i]
  double* const red = sm + 2 * win;        // 3 x 32 warp partials (16-byte aligned: 2*win is even)
  double* const cred = red + 96;           // 3 x 8 CTA partials
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r0 = tid * K;                                         // local
  const int rc0 = r0 < nloc ? r0 : nloc;                          // dead threads sit on the zero guard
  const int nlive = r0 < nloc ? (nloc - r0 < K ? nloc - r0 : K) : 0;
  // the same window of d in the neighbours' shared memory (distributed shared memory)
  auto remote_dv = [&](int rank) -> CheckedPtr<double> {
    double* const rsm = cluster.map_shared_rank(sm, rank);
    return checked(rsm + win + pad, rsm + win, rsm + 2 * win);
  };
  const CheckedPtr<double> dv_prev = crank > 0 ? remote_dv(crank - 1) : checked<double>(nullptr, nullptr, nullptr);
  const CheckedPtr<double> dv_next = crank + 1 < CL ? remote_dv(crank + 1) : checked<double>(nullptr, nullptr, nullptr);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
        static_cast<uint32_t>(__cvta_generic_to_shared(&s_tmem))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  for (int i = tid; i < win; i += BT) dv[i - pad] = 0.0;  // entries nobody writes stay zero (guards past n)
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  cluster.sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tbase =
      s_tmem + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + static_cast<uint32_t>(warp >> 2) * static_cast<uint32_t>(CW);
  [[maybe_unused]] const uint32_t tm_lo = tbase & 0xffffu, tm_hi = tm_lo + CW;

  // own entries of d -> shared memory, border entries also into the neighbours' halos
  auto publish = [&](const double (&dl)[K]) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (j < nlive) {
        const int i = r0 + j;
        dv[i] = dl[j];
        if (dv_prev && i < pad) dv_prev[RC + i] = dl[j];
        if (dv_next && i >= nloc - pad) dv_next[i - nloc] = dl[j];
      }
    }
  };

  auto spmv_strip = [&](const double (&dl)[K], double (&hh)[K]) {
    double ra = dv[rc0 + m - 1], rb = dv[rc0 + m];
    double la = dv[rc0 - m - 1], lb = dv[rc0 - m];
    double below = dv[rc0 - 1];
    double m1_prev = m1[rc0 - 1];
    TmemRow q[2];
    MSGPU_CHECK_TMEM(tbase, 16u, tm_lo, tm_hi);
    tmem_issue_row(tbase, q[0]);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int r = rc0 + j;
      TmemRow& cur = q[j & 1];
      tmem_wait_row(cur);
      if (j + 1 < K) {
        MSGPU_CHECK_TMEM(tbase + 14u * (j + 1), 16u, tm_lo, tm_hi);
        tmem_issue_row(tbase + 14u * (j + 1), q[(j + 1) & 1]);
      }
      const double rc = dv[r + m + 1], lc = dv[r - m + 1];
      const double m1r = m1[r];
      const double above = (j + 1 < K) ? dl[j + 1] : dv[r + 1];
      double s = tmem_get(cur, 0) * dl[j];
      double t = tmem_get(cur, 1) * ra;
      s += m1r * above;
      t += tmem_get(cur, 4) * lc;
      s += m1_prev * below;
      t += tmem_get(cur, 2) * rb;
      s += tmem_get(cur, 5) * lb;
      t += tmem_get(cur, 3) * rc;
      s += tmem_get(cur, 6) * la;
      hh[j] = s + t;
      ra = rb; rb = rc;
      la = lb; lb = lc;
      below = dl[j];
      m1_prev = m1r;
    }
  };

  // Sum over the whole cluster, identical in every thread of every CTA.  `buf` in {0,1,2}.
  auto cluster_sum = [&](double part, int buf) -> double {
    const double mine = block_sum_tree<BT>(part, red + 32 * buf);
    if (tid < CL) cluster.map_shared_rank(cred + 8 * buf, tid)[crank] = mine;
    cluster.sync();
    double t = 0.0;
    for (int c = 0; c < CL; ++c) t += cred[8 * buf + c];
    return t;
  };

  for (;;) {
    cluster.sync();  // every CTA is done with the previous system (its halos, cred, s_k)
    if (crank == 0 && tid == 0) {
      const int k = atomicAdd(a.work_counter, 1);
      for (int c = 0; c < CL; ++c) *cluster.map_shared_rank(&s_k, c) = k;
    }
    cluster.sync();
    const int k = s_k;
    if (k >= a.N) break;

    const double* dg = a.diag + static_cast<size_t>(k) * a.diag_stride;
    for (int i = tid - pad; i < RC + pad; i += BT) {
      const int g = lo + i;
      m1[i] = (g >= 0 && g < n) ? __ldg(dg + ld + g) : 0.0;
    }
    double diag_own[JACOBI ? K : 1];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const bool in = j < nlive;
      const int g = lo + r0 + j;
      double v[7];
      v[0] = in ? __ldg(dg + g) : 0.0;
      v[1] = in ? __ldg(dg + 2 * static_cast<size_t>(ld) + g) : 0.0;
      v[2] = in ? __ldg(dg + 3 * static_cast<size_t>(ld) + g) : 0.0;
      v[3] = in ? __ldg(dg + 4 * static_cast<size_t>(ld) + g) : 0.0;
      v[4] = (in && g - (m - 1) >= 0) ? __ldg(dg + 2 * static_cast<size_t>(ld) + g - (m - 1)) : 0.0;
      v[5] = (in && g - m >= 0) ? __ldg(dg + 3 * static_cast<size_t>(ld) + g - m) : 0.0;
      v[6] = (in && g - (m + 1) >= 0) ? __ldg(dg + 4 * static_cast<size_t>(ld) + g - (m + 1)) : 0.0;
      if (JACOBI) diag_own[j] = in ? v[0] : 1.0;
      MSGPU_CHECK_TMEM(tbase + 14u * j, 16u, tm_lo, tm_hi);
      tmem_store_row(tbase + 14u * j, v);
    }

    double xx[K], gg[K], dl[K], hh[K];
    const double* bk = a.b + static_cast<size_t>(k) * n;
    double* xk = a.x + static_cast<size_t>(k) * n;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      xx[j] = (j < nlive) ? xk[lo + r0 + j] : 0.0;
      dl[j] = xx[j];
    }
    publish(dl);
    cluster.sync();
    spmv_strip(dl, hh);
    double part = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      gg[j] = (j < nlive) ? (hh[j] - bk[lo + r0 + j]) : 0.0;
      part += gg[j] * gg[j];
    }
    double res = sqrt(cluster_sum(part, 0));
    int it = 0;
    bool ok = res <= a.tol;
    double gh = 0.0;
    if (!ok) {
      if (JACOBI) {
        part = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const double z = gg[j] / diag_own[j];
          dl[j] = -z;
          part += gg[j] * z;
        }
        gh = cluster_sum(part, 1);
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) dl[j] = -gg[j];
        gh = res * res;
      }
      publish(dl);
      for (;;) {
        ++it;
        cluster.sync();
        spmv_strip(dl, hh);
        part = strip_dot<K, NACC>([&](int j) { return dl[j] * hh[j]; });
        const double alpha = gh / cluster_sum(part, 0);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          xx[j] += alpha * dl[j];
          gg[j] += alpha * hh[j];
        }
        part = strip_dot<K, NACC>([&](int j) { return gg[j] * gg[j]; });
        res = sqrt(cluster_sum(part, 1));
        if (res <= a.tol) {
          ok = true;
          break;
        }
        if (it >= a.max_it || res != res) break;
        double beta = gh;
        if (JACOBI) {
          part = 0.0;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            hh[j] = gg[j] / diag_own[j];
            part += gg[j] * hh[j];
          }
          gh = cluster_sum(part, 2);
          beta = gh / beta;
#pragma unroll
          for (int j = 0; j < K; ++j) dl[j] = beta * dl[j] - hh[j];
        } else {
          gh = res * res;
          beta = gh / beta;
#pragma unroll
          for (int j = 0; j < K; ++j) dl[j] = beta * dl[j] - gg[j];
        }
        publish(dl);
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (j < nlive) xk[lo + r0 + j] = xx[j];
    if (crank == 0 && tid == 0) {
      if (a.iters) a.iters[k] = it;
      if (a.res) a.res[k] = res;
      if (!ok) atomicExch(a.fail_flag, 1);
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(s_tmem));
  cluster.sync();  // no CTA may exit while a neighbour could still address its shared memory
}

// -------------------------------------------------------------------------------------------
// Global-memory CG for systems that do not fit in one SM's shared memory (m > 67): the matrix is
// streamed from L2/HBM every iteration, work vectors live in a per-CTA workspace.  Same
// recurrence, same reduction order; kept simple — the cluster/DSMEM variant is the next step.
template <int BT>
__global__ void __launch_bounds__(BT) k_cg_global(CgArgs a, double* __restrict__ workspace) {
  if (a.gate != nullptr && ((*a.gate == 0) != (a.gate_run_if_zero != 0))) return;  // the other CG kernel does the work
  __shared__ double red[96];
  __shared__ int s_k;
  const int n = a.md.n, m = a.md.m, ld = a.md.ld;
  const int tid = threadIdx.x;
  double* g = workspace + static_cast<size_t>(blockIdx.x) * 3 * n;
  double* d = g + n;
  double* h = d + n;
  const int off[4] = {1, m - 1, m, m + 1};
  for (;;) {
    __syncthreads();
    if (tid == 0) s_k = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int k = s_k;
    if (k >= a.N) break;
    const double* dg = a.diag + static_cast<size_t>(k) * a.diag_stride;
    const double* bk = a.b + static_cast<size_t>(k) * n;
    double* xk = a.x + static_cast<size_t>(k) * n;
    auto spmv = [&](const double* v, double* out) {
      for (int r = tid; r < n; r += BT) {
        double s = dg[r] * v[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double* dj = dg + static_cast<size_t>(j + 1) * ld;
          if (r + off[j] < n) s += dj[r] * v[r + off[j]];
          if (r - off[j] >= 0) s += dj[r - off[j]] * v[r - off[j]];
        }
        out[r] = s;
      }
    };
    spmv(xk, h);
    double part = 0.0;
    for (int r = tid; r < n; r += BT) {
      const double gr = h[r] - bk[r];
      g[r] = gr;
      part += gr * gr;
    }
    double res = sqrt(block_sum<BT>(part, red));
    int it = 0;
    bool ok = res <= a.tol;
    if (!ok) {
      double gh;
      if (a.precond == 1) {
        part = 0.0;
        for (int r = tid; r < n; r += BT) {
          const double z = g[r] / dg[r];
          d[r] = -z;
          part += g[r] * z;
        }
        gh = block_sum<BT>(part, red + 32);
      } else {
        for (int r = tid; r < n; r += BT) d[r] = -g[r];
        gh = res * res;
      }
      for (;;) {
        ++it;
        __syncthreads();
        spmv(d, h);
        part = 0.0;
        for (int r = tid; r < n; r += BT) part += d[r] * h[r];
        const double alpha = gh / block_sum<BT>(part, red);
        part = 0.0;
        for (int r = tid; r < n; r += BT) {
          xk[r] += alpha * d[r];
          const double gr = g[r] + alpha * h[r];
          g[r] = gr;
          part += gr * gr;
        }
        res = sqrt(block_sum<BT>(part, red + 32));
        if (res <= a.tol) {
          ok = true;
          break;
        }
        if (it >= a.max_it || res != res) break;
        double beta = gh;
        if (a.precond == 1) {
          part = 0.0;
          for (int r = tid; r < n; r += BT) {
            const double z = g[r] / dg[r];
            h[r] = z;
            part += g[r] * z;
          }
          gh = block_sum<BT>(part, red + 64);
          beta = gh / beta;
          for (int r = tid; r < n; r += BT) d[r] = beta * d[r] - h[r];
        } else {
          gh = res * res;
          beta = gh / beta;
          for (int r = tid; r < n; r += BT) d[r] = beta * d[r] - g[r];
        }
      }
    }
    if (tid == 0) {
      if (a.iters) a.iters[k] = it;
      if (a.res) a.res[k] = res;
      if (!ok) atomicExch(a.fail_flag, 1);
    }
  }
}

// -------------------------------------------------------------------------------------------
// K5/K6: one warp per system, segmented warp reduction.
//   elliptic  c0 = sum_r x[r] * jw[r]                      = int v_k det(F) dy
//   parabolic c0 = sum_r x[r] * jw[r] (shared lumped mass) = int rho_k dy
//   bio       c0 = sum_left  (-kappa_2 x jL + bcL),  c1 = sum_right (-kappa_4 x jR + bcR)
__global__ void __launch_bounds__(128) k_coupling(CouplingArgs a) {
  const int warp = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int k = warp;
  const MeshDims& md = a.md;
  if (k < a.N) {
    const double* x = a.x + static_cast<size_t>(k) * md.n;
    if (a.problem == PF_BIO) {
      const double* edge = a.edge + static_cast<size_t>(k) * 4 * md.m;
      double su = 0.0, sw = 0.0;
      for (int iy = lane; iy < md.m; iy += 32) {
        su += -a.kappa_left * x[iy] * edge[iy] + edge[2 * md.m + iy];
        sw += -a.kappa_right * x[(md.m - 1) * md.m + iy] * edge[md.m + iy] + edge[3 * md.m + iy];
      }
      su = warp_sum(su);
      sw = warp_sum(sw);
      if (lane == 0) {
        a.c0[k] = su;
        a.c1[k] = sw;
        for (int p = 0; p < a.n_peers; ++p) {
          a.peer_c0[p][k] = su;
          a.peer_c1[p][k] = sw;
        }
      }
    } else {
      const double* jw = a.jw + static_cast<size_t>(k) * a.jw_stride;
      double s = 0.0;
      for (int r = lane; r < md.n; r += 32) s += x[r] * jw[r];
      s = warp_sum(s);
      if (lane == 0) {
        a.c0[k] = s;
        for (int p = 0; p < a.n_peers; ++p) a.peer_c0[p][k] = s;
      }
    }
  }
  if (a.n_peers > 0) {
    // signal: all peer stores of this CTA are performed system-wide, then the last CTA raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int done = atomicAdd(a.done_counter, 1);
      if (done == static_cast<int>(gridDim.x) - 1) {
        __threadfence_system();
        for (int p = 0; p < a.n_peers; ++p) *reinterpret_cast<volatile unsigned int*>(a.peer_flag[p]) = a.flag_value;
        *a.done_counter = 0;
      }
    }
  }
}

__global__ void __launch_bounds__(128) k_shared_spmv(MeshDims md, int N, const double* __restrict__ dg,
                                                      const double* __restrict__ x, double* __restrict__ y) {
  // blockIdx.x: system, blockIdx.y: block of 128 nodes
  const int k = static_cast<int>(blockIdx.x), r = static_cast<int>(blockIdx.y) * 128 + static_cast<int>(threadIdx.x);
  if (k >= N || r >= md.n) return;
  const size_t t = static_cast<size_t>(k) * md.n + r;
  const double* v = x + static_cast<size_t>(k) * md.n;
  const int off[4] = {1, md.m - 1, md.m, md.m + 1};
  double s = dg[r] * v[r];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double* dj = dg + static_cast<size_t>(j + 1) * md.ld;
    if (r + off[j] < md.n) s += dj[r] * v[r + off[j]];
    if (r - off[j] >= 0) s += dj[r - off[j]] * v[r - off[j]];
  }
  y[t] = s;
}

size_t resident_smem_bytes(const MeshDims& md) {
  const int pad = (md.m + 2) & ~1;
  return (static_cast<size_t>(NDIAG) * (pad + md.ld) + md.n + 2 * pad + 96) * sizeof(double);
}

template <int BT, int K>
int launch_strip(cudaStream_t s, const CgArgs& a, int num_sms) {
  const size_t smem = resident_smem_bytes(a.md);
  auto kern = a.precond == 2 ? k_cg_strip<BT, K, 2> : a.precond == 1 ? k_cg_strip<BT, K, 1> : k_cg_strip<BT, K, 0>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem);
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(num_sms) * per_sm;
  if (grid > a.N) grid = a.N;
  kern<<<static_cast<int>(grid), BT, smem, s>>>(a);
  return 1;
}

size_t tmem_smem_bytes(const MeshDims& md, int K) {
  const int pad = tmem_pad(md.m, K);
  return (static_cast<size_t>(4) * (pad + md.ld) + md.n + 2 * pad + 98) * sizeof(double);
}

template <int BT, int K>
int launch_strip_tmem(cudaStream_t s, const CgArgs& a, int num_sms, bool debug) {
  const size_t smem = tmem_smem_bytes(a.md, K);
  auto kern = a.precond == 1 ? k_cg_strip_tmem<BT, K, true> : k_cg_strip_tmem<BT, K, false>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  // CTAs per SM: bounded by the TMEM columns a CTA allocates (not known to the occupancy calculator), by shared
  // memory and by registers
  int per_sm = 1;
  const int by_tmem = 512 / tmem_alloc_columns(BT, K);
  // ask for the largest shared-memory carve-out: with the default preference the calculator (and the launch) may
  // settle for a split that holds a single CTA of this size
  if (by_tmem > 1) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const cudaError_t occ = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem);
  if (debug)
    std::fprintf(stderr, "[msgpu] k_cg_strip_tmem<%d,%d>: occupancy query %s -> %d CTAs per SM, TMEM allows %d, smem %zu B\n", BT, K,
                 cudaGetErrorName(occ), per_sm, by_tmem, smem);
  // The calculator answers 1 for the 256 x 9 shape although two CTAs fit (2 x 84.5 KB of shared memory, 2 x 32 K
  // registers — ncu reports both limits as 2).  The kernel is persistent with an atomic work queue, so offering the
  // hardware by_tmem CTAs per SM is safe either way: a CTA that only becomes resident late finds the queue drained.
  cudaFuncAttributes fa{};
  if (by_tmem > per_sm && cudaFuncGetAttributes(&fa, kern) == cudaSuccess &&
      static_cast<size_t>(by_tmem) * (smem + fa.sharedSizeBytes + 1024) <= 227u * 1024u &&
      static_cast<long long>(by_tmem) * BT * fa.numRegs <= 65536)
    per_sm = by_tmem;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(num_sms) * per_sm;
  if (grid > a.N) grid = a.N;
  kern<<<static_cast<int>(grid), BT, smem, s>>>(a);
  return 1;
}

// Cluster variant: CL = ceil(n / (256*17)) CTAs per system, as many clusters as the GPU can hold.
int launch_cluster(cudaStream_t s, const CgArgs& a, int num_sms) {
  constexpr int K = 17, RC = 256 * K;
  const int CL = (a.md.n + RC - 1) / RC;
  if (CL < 1 || CL > 8) return -1;
  const int pad = tmem_pad(a.md.m, K);
  const size_t smem = (static_cast<size_t>(2) * (RC + 2 * pad) + 96 + 24) * sizeof(double);
  auto kern = a.precond == 1 ? k_cg_cluster<K, true> : k_cg_cluster<K, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return -1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(CL);
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess || clusters < 1) {
    cudaGetLastError();
    clusters = num_sms / CL / 2;
    if (clusters < 1) clusters = 1;
  }
  if (clusters > a.N) clusters = a.N;
  cfg.gridDim = dim3(static_cast<unsigned>(clusters) * CL);
  CgArgs args = a;
  if (cudaLaunchKernelEx(&cfg, kern, args) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return 1;
}

#ifdef MSGPU_AB_VARIANTS
template <int BT, int RPT>
int launch_resident(cudaStream_t s, const CgArgs& a, int num_sms) {
  const size_t smem = resident_smem_bytes(a.md);
  auto kern = a.precond == 1 ? k_cg_resident<BT, RPT, true> : k_cg_resident<BT, RPT, false>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem);
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(num_sms) * per_sm;
  if (grid > a.N) grid = a.N;
  kern<<<static_cast<int>(grid), BT, smem, s>>>(a);
  return 1;
}

#endif  // MSGPU_AB_VARIANTS

}  // namespace

CgSwitches cg_switches_from_env() {
  CgSwitches w;
  auto flag = [](const char* name, char c) {
    const char* e = std::getenv(name);
    return e != nullptr && e[0] == c;
  };
  w.stencil = flag("MSGPU_CG_STENCIL", '0') ? 0 : 1;
  w.stencil_tmem = flag("MSGPU_STENCIL_TMEM", '0') ? 0 : 1;
  if (const char* e = std::getenv("MSGPU_STENCIL_CTAS")) w.stencil_ctas = std::atoi(e);
  if (const char* e = std::getenv("MSGPU_STENCIL_VARIANT")) w.stencil_variant = std::atoi(e);
  w.force_global = flag("MSGPU_CG_VARIANT", 'g') ? 1 : 0;
  w.wide = flag("MSGPU_CG_WIDE", '1') ? 1 : 0;
  w.strided = flag("MSGPU_CG_STRIDED", '1') ? 1 : 0;
  w.tmem = flag("MSGPU_CG_TMEM", '0') ? 0 : (flag("MSGPU_CG_TMEM", '5') ? 5 : -1);
  w.tmem_small = flag("MSGPU_CG_TMEM_SMALL", '0') ? 0 : 1;
  w.cluster = flag("MSGPU_CG_CLUSTER", '0') ? 0 : 1;
  w.debug_launch = std::getenv("MSGPU_DEBUG_LAUNCH") != nullptr ? 1 : 0;
  if (const char* e = std::getenv("MSGPU_SSOR_OMEGA")) {
    const double o = std::atof(e);
    if (o > 0.0 && o < 2.0) w.ssor_omega = o;
  }
  return w;
}

static double squared_threshold(double tol) {
  if (!(tol > 0.0)) return tol == 0.0 ? 0.0 : -1.0;
  double t = tol * tol;
  while (std::sqrt(std::nextafter(t, INFINITY)) <= tol) t = std::nextafter(t, INFINITY);
  while (std::sqrt(t) > tol) t = std::nextafter(t, 0.0);
  return t;
}

static int launch_cg_general(cudaStream_t s, const CgArgs& a_in, int num_sms, int* variant, const CgSwitches& sw) {
  if (a_in.precond == 2 && !(resident_smem_bytes(a_in.md) <= 227 * 1024 && a_in.md.n <= 512 * 9)) return -2;
  CgArgs a = a_in;
  const int n = a.md.n;
  const bool fits = resident_smem_bytes(a.md) <= 227 * 1024 && n <= 512 * 9;  // 512 x 9 rows is the largest strip shape
  if (fits && !sw.force_global) {
    if (variant) *variant = 0;
    // default for the large resident sizes: matrix in tensor memory (tmem == 0: plain strip kernel, A/B); the SSOR
    // sweeps need all five diagonals in shared memory and live in the plain strip kernel
    if (a.precond == 2) {
      if (variant) *variant = 2;
      if (n <= 32) return launch_strip<32, 1>(s, a, num_sms);
      if (n <= 128) return launch_strip<128, 1>(s, a, num_sms);
      if (n <= 384) return launch_strip<128, 3>(s, a, num_sms);
      if (n <= 1280) return launch_strip<256, 5>(s, a, num_sms);
      if (n <= 2560) return launch_strip<512, 5>(s, a, num_sms);
      return launch_strip<512, 9>(s, a, num_sms);
    }
    if (sw.tmem != 0 && n > 2560) {
      if (variant) *variant = 3;
      if (n <= 256 * 17 && sw.tmem != 5) return launch_strip_tmem<256, 17>(s, a, num_sms, sw.debug_launch != 0);
      if (n <= 512 * 9) return launch_strip_tmem<512, 9>(s, a, num_sms, sw.debug_launch != 0);
    }
    // mid sizes (1153..2304 rows, e.g. 41 x 41 nodes): 256 threads x 9 rows allocate 256 TMEM columns, two systems
    // share an SM (3.0 instead of 4.4 ms against the plain strip kernel at 2401 systems of 2025 rows).  Below that
    // the plain strip kernel with 3-4 CTAs per SM is the faster one (measured: 128 x 9 with 128 columns loses 12 %).
    if (sw.tmem != 0 && sw.tmem_small && n > 128 * 9 && n <= 256 * 9) {
      if (variant) *variant = 3;
      return launch_strip_tmem<256, 9>(s, a, num_sms, sw.debug_launch != 0);
    }
#ifdef MSGPU_AB_VARIANTS
    if (sw.strided) {  // first generation: strided row ownership, everything in shared memory
      if (n <= 32) return launch_resident<32, 1>(s, a, num_sms);
      if (n <= 128) return launch_resident<128, 1>(s, a, num_sms);
      if (n <= 384) return launch_resident<128, 3>(s, a, num_sms);
      if (n <= 1280) return launch_resident<256, 5>(s, a, num_sms);
      if (n <= 2560) return launch_resident<512, 5>(s, a, num_sms);
      if (sw.wide) return launch_resident<1024, 5>(s, a, num_sms);
      return launch_resident<512, 9>(s, a, num_sms);
    }
#endif
    if (variant) *variant = 2;
    if (n <= 32) return launch_strip<32, 1>(s, a, num_sms);
    if (n <= 128) return launch_strip<128, 1>(s, a, num_sms);
    if (n <= 384) return launch_strip<128, 3>(s, a, num_sms);
    if (n <= 1280) return launch_strip<256, 5>(s, a, num_sms);
    if (n <= 2560) return launch_strip<512, 5>(s, a, num_sms);
    return launch_strip<512, 9>(s, a, num_sms);
  }
  if (!sw.force_global && sw.cluster && n > 2560 && n <= 8 * 256 * 17) {
    const int r = launch_cluster(s, a, num_sms);
    if (r > 0) {
      if (variant) *variant = 4;
      return r;
    }
  }
  if (variant) *variant = 1;
  constexpr int BT = 512;
  int grid = num_sms * 2;
  if (grid > a.N) grid = a.N;
  const size_t need = static_cast<size_t>(grid) * 3 * n * sizeof(double);
  if (a.workspace == nullptr || a.workspace_bytes == nullptr) return -1;
  if (need > *a.workspace_bytes) {
    if (*a.workspace) {
      cudaStreamSynchronize(s);  // an earlier launch on this stream may still use the old buffer
      cudaFree(*a.workspace);
      *a.workspace = nullptr;
      *a.workspace_bytes = 0;
    }
    if (cudaMalloc(a.workspace, need) != cudaSuccess) return -1;
    *a.workspace_bytes = need;
  }
  k_cg_global<BT><<<grid, BT, 0, s>>>(a, *a.workspace);
  return 1;
}

// Dispatch.  Systems whose matrix is class-uniform (kernels_cg_stencil.cu) take the stencil kernel; whether they
// are is decided ON THE DEVICE by the compression kernel unless the caller already knows (stencil_hint): both CG
// kernels are launched back to back behind the flag and one of them returns at once.
int launch_cg(cudaStream_t s, const CgArgs& a_in, int num_sms, int* variant) {
  if (a_in.N == 0) return 0;
  if (a_in.precond == 2 && !(resident_smem_bytes(a_in.md) <= 227 * 1024 && a_in.md.n <= 512 * 9)) return -2;
  CgArgs a = a_in;
  a.tol2 = squared_threshold(a.tol);
  a.gate = nullptr;
  cudaMemsetAsync(a.work_counter, 0, sizeof(int), s);
  cudaMemsetAsync(a.fail_flag, 0, sizeof(int), s);
  int launched = 0, stencil_variant = 0;
  const CgSwitches sw = a.sw ? *a.sw : cg_switches_from_env();
  a.sw = nullptr;  // host pointer: not for the kernels
  if (a.stencil_ws != nullptr && a.stencil_bad != nullptr && a.precond == 0 && a.stencil_hint != 2 && sw.stencil &&
      stencil_supported(a.md, a.dirichlet)) {
    cudaMemsetAsync(a.stencil_bad, 0, sizeof(int), s);
    launched += launch_stencil_compress(s, a, a.stencil_hint == 0);
    CgArgs st = a;
    st.gate = a.stencil_bad;
    st.gate_run_if_zero = 1;
    const int r = launch_cg_stencil(s, st, num_sms, sw);
    if (r > 0) {
      launched += r;
      stencil_variant = 5;
      if (a.stencil_hint == 1) {
        if (variant) *variant = 5;
        return launched;
      }
      a.gate = a.stencil_bad;  // the general kernel below runs only if the flag went up
      a.gate_run_if_zero = 0;
    }
  }
  if (a.x_in != nullptr && a.x_in != a.x)  // the general kernels solve in place: bring the start vectors over
    cudaMemcpyAsync(a.x, a.x_in, sizeof(double) * static_cast<size_t>(a.N) * a.md.n, cudaMemcpyDeviceToDevice, s);
  int general_variant = 0;
  const int r = launch_cg_general(s, a, num_sms, &general_variant, sw);
  if (r < 0) return r;
  if (variant) *variant = stencil_variant ? (stencil_variant | (general_variant << 8)) : general_variant;
  return launched + r;
}

int launch_coupling(cudaStream_t s, const CouplingArgs& a) {
  if (a.N == 0) return 0;
  const int blocks = (a.N * 32 + 127) / 128;
  k_coupling<<<blocks, 128, 0, s>>>(a);
  return 1;
}

int launch_shared_spmv(cudaStream_t s, MeshDims md, int N, const double* diag, const double* x, double* y) {
  if (N == 0) return 0;
  const dim3 grid(static_cast<unsigned>(N), static_cast<unsigned>((md.n + 127) / 128));
  k_shared_spmv<<<grid, 128, 0, s>>>(md, N, diag, x, y);
  return 1;
}

}  // namespace msgpu
