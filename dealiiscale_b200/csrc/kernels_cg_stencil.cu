// Batched FP64 conjugate gradients for CLASS-UNIFORM systems on sm_100a (kernel K4s).
//
// Same job and same recurrence as k_cg_strip_tmem (kernels_cg.cu): deal.II 9.1's SolverCG with
// PreconditionIdentity, once per micro system (reference: src/elliptic-elliptic/micro.cpp:175-182,
// src/bio-multiscale/micro.cpp:309-314, src/elliptic-parabolic/rho_solver.cpp:222-229).
//
// What is different.  When the pulled-back coefficient K det(F) of a micro problem does not vary over the micro
// mesh (F constant per system: every bio parameter set, the linear / scaled / rotated elliptic maps, the parabolic
// driver), every cell contributes the same 4 x 4 block, so the assembled matrix has only NINE distinct rows - one
// per position class {first, inner, last} x {first, inner, last} of the node - and for Dirichlet problems, once the
// eliminated boundary rows are left out, a single one.  k_stencil_compress extracts the 9 x 9 class table of every
// system from the assembled diagonals and (unless the caller vouches for an unchanged matrix) checks EVERY stored
// entry against it; a device flag then decides between this kernel and the general one without a host round trip.
// The entries are the assembled doubles themselves: the matrix-vector product is the one with the explicit matrix
// (only the order of the nine additions of a row and of the reduction trees differs from k_cg_strip_tmem).
//
// With the matrix gone from tensor memory the SM has room for TWO systems, which is what the general kernel lacks
// (45 % of a CG step is two block reductions and a scalar chain with nothing else resident to hide them):
//   * a thread owns a strip of K consecutive nodes of ONE mesh column (S strips per column, S*K >= nodes per axis;
//     the last strip is anchored at the end of the column, so both boundary nodes of a column sit at compile-time
//     positions - row 0 of the first strip, row K-1 of the last - and the S*K - mp rows the last strip shares with
//     its neighbour are computed twice, bit-identically, and counted once);
//     the search direction d sits in registers (own strip) and in shared memory (neighbours: column pitch P with a
//     zero pad that doubles as the missing neighbours of boundary nodes, so there are no index tests);
//   * x, g and h = A d live in TENSOR MEMORY, 6K columns per thread, streamed through registers four rows at a time
//     (tcgen05.ld / tcgen05.st, SASS LDTM / STTM) - lane-private storage behind its own datapath;
//   * the nine couplings of the thread's column class are registers; the two boundary rows of a column (first row
//     of the first strip, last live row of the last strip) take theirs from the class table in shared memory;
//   * <= 88 registers x 352 threads and 256 tensor-memory columns per CTA: two CTAs per SM.
// Thread t -> (strip s = t / mp, column ix = t % mp): a warp holds strips of the same s (but for one seam), so the
// "does this warp hold a boundary row" tests are warp-uniform and nothing diverges around the .aligned tcgen05 ops.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "cg_common.cuh"
#include "kernels.h"

namespace msgpu {

namespace {

using cgc::l2_prefetch_bulk;
using cgc::warp_sum;

constexpr int ST_MAX_THREADS = 352;  // 11 warps; two CTAs of 88 registers per thread fill the register file
constexpr int ST_RED = 16;           // slots of one reduction buffer (>= warps per CTA; a power of two)

struct StencilPlan {
  int mp;     // nodes per axis of the block the CG runs on: m, or m - 2 when Dirichlet rows are left out
  int off;    // first mesh line of that block (0 | 1)
  int K;      // rows per thread
  int S;      // strips per mesh column
  int P;      // pitch of a column of d in shared memory (odd, >= mp + 1)
  int ov;     // rows the last strip of a column shares with the strip before it (S*K - mp <= 3)
  int block;  // CTA size
  int alloc;  // tensor-memory columns per CTA (power of two)
  int cw;     // tensor-memory columns per warp
};

bool make_plan(const MeshDims& md, int dirichlet, StencilPlan& best) {
  const int off = dirichlet ? 1 : 0, mp = md.m - 2 * off;
  if (mp < 24) return false;  // small systems: several CTAs of the plain strip kernel per SM do as well
  double best_eff = 0.0;
  for (int K = 13; K >= 9; --K) {
    const int S = (mp + K - 1) / K, ov = S * K - mp, threads = S * mp;
    if (ov > 3 || threads > ST_MAX_THREADS || S < 2) continue;
    const int block = (threads + 31) & ~31;
    const double eff = static_cast<double>(mp) * mp / (static_cast<double>(block) * K);
    if (eff <= best_eff) continue;
    StencilPlan p;
    p.mp = mp;
    p.off = off;
    p.K = K;
    p.S = S;
    p.P = (mp + 1) | 1;
    p.ov = ov;
    p.block = block;
    p.cw = 6 * K;
    const int slots = (block / 32 + 3) / 4;
    int alloc = 32;
    while (alloc < slots * p.cw) alloc *= 2;
    if (alloc > 512) continue;
    p.alloc = alloc;
    best = p;
    best_eff = eff;
  }
  return best_eff > 0.0;
}

size_t plan_smem_bytes(const StencilPlan& p) {
  const size_t dvlen = (static_cast<size_t>(2) + static_cast<size_t>(p.mp + 4) * p.P + 1) & ~static_cast<size_t>(1);
  return (dvlen + 3 * ST_RED + STENCIL_TABLE + 1) * sizeof(double);
}

// ------------------------------------------------------------------------------------------------
// tensor-memory vectors: ND doubles = 2 ND columns of one lane
template <int ND>
struct TV {
  uint32_t r[2 * ND];
};
template <int ND>
__device__ __forceinline__ double tv_get(const TV<ND>& q, int i) {
  return __hiloint2double(static_cast<int>(q.r[2 * i + 1]), static_cast<int>(q.r[2 * i]));
}

// Asynchronous loads; tm_wait* takes the destination registers as in/out operands so that no use of them can be
// scheduled before the wait (tensor memory is not aliased by any pointer: no "memory" clobber needed).
template <int ND>
__device__ __forceinline__ void tm_ld(uint32_t a, TV<ND>& q) {
  if constexpr (ND == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]), "=r"(q.r[4]), "=r"(q.r[5]), "=r"(q.r[6]),
                   "=r"(q.r[7])
                 : "r"(a));
  } else if constexpr (ND == 2) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3])
                 : "r"(a));
  } else {
    static_assert(ND == 1, "chunk size");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(q.r[0]), "=r"(q.r[1]) : "r"(a));
  }
}
template <int ND>
__device__ __forceinline__ void tm_wait(TV<ND>& q) {
  if constexpr (ND == 4) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(q.r[0]), "+r"(q.r[1]), "+r"(q.r[2]), "+r"(q.r[3]), "+r"(q.r[4]), "+r"(q.r[5]), "+r"(q.r[6]),
                   "+r"(q.r[7]));
  } else if constexpr (ND == 2) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(q.r[0]), "+r"(q.r[1]), "+r"(q.r[2]), "+r"(q.r[3]));
  } else {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(q.r[0]), "+r"(q.r[1]));
  }
}
template <int ND>
__device__ __forceinline__ void tm_wait3(TV<ND>& a, TV<ND>& b, TV<ND>& c) {
  if constexpr (ND == 4) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(a.r[4]), "+r"(a.r[5]), "+r"(a.r[6]),
                   "+r"(a.r[7]), "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]),
                   "+r"(b.r[6]), "+r"(b.r[7]), "+r"(c.r[0]), "+r"(c.r[1]), "+r"(c.r[2]), "+r"(c.r[3]), "+r"(c.r[4]),
                   "+r"(c.r[5]), "+r"(c.r[6]), "+r"(c.r[7]));
  } else if constexpr (ND == 2) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]),
                   "+r"(b.r[3]), "+r"(c.r[0]), "+r"(c.r[1]), "+r"(c.r[2]), "+r"(c.r[3]));
  } else {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(b.r[0]), "+r"(b.r[1]), "+r"(c.r[0]), "+r"(c.r[1]));
  }
}
// Asynchronous store of ONE double (two columns); tm_st_wait before the same thread reads the columns back.  Wider
// stores want their 4 / 8 source registers consecutive and aligned, which costs more register moves than the extra
// store instructions do (ncu, round 2: 190 of 680 instructions per CG step were such moves).
__device__ __forceinline__ void tm_st1(uint32_t a, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(a), "r"(__double2loint(v)),
               "r"(__double2hiint(v)));
}
__device__ __forceinline__ void tm_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n"); }

// f(integral_constant<int, ND>, j0) for the chunks [j0, j0 + ND) of a strip of K rows: fours, then a pair, then one
template <int K, class F>
__device__ __forceinline__ void for_each_chunk(F&& f) {
#pragma unroll
  for (int c = 0; c < K / 4; ++c) f(std::integral_constant<int, 4>{}, 4 * c);
  if constexpr ((K % 4) >= 2) f(std::integral_constant<int, 2>{}, K - (K % 4));
  if constexpr ((K % 2) == 1) f(std::integral_constant<int, 1>{}, K - 1);
}

// Sum over the CTA, result in every thread: warp partials into `red` (ST_RED slots; the slots of warps the CTA does
// not have stay zero), one barrier, then every warp adds them with one more shuffle tree (fixed order:
// bit-reproducible, and 4 additions per thread instead of 11 - the FP64 pipe is what this kernel runs out of).
// `red` must not be written again before the next barrier that follows this call (callers rotate three buffers).
__device__ __forceinline__ double block_sum12(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31;
  if (lane == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = red[lane & 15];  // slots 12..15 are zero padding
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// One matrix row: nine products, one chain, fixed order (a thread has K independent rows in flight, and an extra
// addition per row to join two chains is FP64 work this kernel cannot afford).  c: self, up, down, right-down, right,
// right-up, left-up, left, left-down.
__device__ __forceinline__ double row9(const double* __restrict__ c, double self, double above, double below, double la,
                                       double lb, double lc, double ra, double rb, double rc) {
  double s = c[0] * self;
  s += c[1] * above;
  s += c[2] * below;
  s += c[3] * ra;
  s += c[6] * lc;
  s += c[4] * rb;
  s += c[7] * lb;
  s += c[5] * rc;
  s += c[8] * la;
  return s;
}

// The same row for an inner node of a column: in a class-uniform symmetric matrix its up / down couplings coincide
// (c1) and so do the three pairs of couplings to the right and left columns (cs[0..2] = right-down | left-up,
// right | left, right-up | left-down) - k_stencil_compress checks exactly that - so five registers carry the row.
// On the first / last column the side that does not exist multiplies the zero guard column.
__device__ __forceinline__ double row5(double c0, double c1, const double (&cs)[3], double self, double above,
                                       double below, double la, double lb, double lc, double ra, double rb, double rc) {
  double s = c0 * self;
  s += c1 * above;
  s += c1 * below;
  s += cs[0] * ra;
  s += cs[0] * lc;
  s += cs[1] * rb;
  s += cs[1] * lb;
  s += cs[2] * rc;
  s += cs[2] * la;
  return s;
}

template <int K>
__global__ void __launch_bounds__(ST_MAX_THREADS, 2)
    k_cg_stencil(const __grid_constant__ CgArgs a, const __grid_constant__ StencilPlan p) {
  if (a.gate != nullptr && ((*a.gate == 0) != (a.gate_run_if_zero != 0))) return;  // the other kernel does the work
  constexpr bool FAST = MSGPU_FAST_SCALARS;
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_k, s_next;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int mp = p.mp, P = p.P, S = p.S, m = a.md.m, n = a.md.n, off = p.off, ov = p.ov;
  const int dvlen = (2 + (mp + 4) * P + 1) & ~1;
  double* const dv = sm + 2 + P;   // dv[ix * P + iy], ix in [-1, mp + 2]; everything nobody writes stays zero
  double* const red = sm + dvlen;  // 3 x ST_RED, 16-byte aligned
  double* const tab = red + 3 * ST_RED;

  const int s = tid / mp;
  const bool live = s < S;  // threads past S*mp idle on column mp + 1: zeros there and on both sides, so h = 0
  const int ix = live ? tid - s * mp : mp + 1;
  const bool first = live && s == 0, last = live && s == S - 1;
  double* const dvc = dv + ix * P + (live ? (last ? mp - K : s * K) : 0);
  const int cx = ix == 0 ? 0 : (ix == mp - 1 ? 2 : 1);
  // class rows of the first / last node of this strip: the inner class unless the strip touches the end of the column
  const double* const tab_lo = tab + (cx * 3 + (first ? 0 : 1)) * 9;
  const double* const tab_hi = tab + (cx * 3 + (last ? 2 : 1)) * 9;
  // rows [0, ov) of a last strip belong to the strip before it: same arithmetic there and here, counted there
  auto shadow = [&](int j) -> bool { return j < 3 && last && j < ov; };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(&s_tmem))),
                 "r"(static_cast<uint32_t>(p.alloc)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  for (int i = tid; i < dvlen + 3 * ST_RED + STENCIL_TABLE + 1; i += blockDim.x) sm[i] = 0.0;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // lane quadrant of this warp in bits 31:16, column block of this warp in bits 15:0
  const uint32_t xcol = s_tmem + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) +
                        static_cast<uint32_t>(warp >> 2) * static_cast<uint32_t>(p.cw);
  const uint32_t gcol = xcol + 2u * K, hcol = xcol + 4u * K;

  double dl[K];      // own strip of the search direction
  double c0, c1;     // couplings of an inner node of this thread's column: self, up = down
  double cs[3];      // ... and to the neighbouring columns (the same for every thread)

  // h = A d for this strip -> tensor memory; returns the strip's part of d.h
  auto spmv = [&]() -> double {
    double la = dvc[-P - 1], lb = dvc[-P];
    double ra = dvc[P - 1], rb = dvc[P];
    double below = dvc[-1];
    double acc[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const double lc = dvc[-P + j + 1], rc = dvc[P + j + 1];
      const double above = (j + 1 < K) ? dl[j + 1] : dvc[K];
      double h;
      if (j == 0)
        h = row9(tab_lo, dl[j], above, below, la, lb, lc, ra, rb, rc);
      else if (j == K - 1)
        h = row9(tab_hi, dl[j], above, below, la, lb, lc, ra, rb, rc);
      else
        h = row5(c0, c1, cs, dl[j], above, below, la, lb, lc, ra, rb, rc);
      acc[j & 1] += dl[j] * (shadow(j) ? 0.0 : h);
      tm_st1(hcol + 2u * j, h);
      la = lb;
      lb = lc;
      ra = rb;
      rb = rc;
      below = dl[j];
    }
    tm_st_wait();
    return acc[0] + acc[1];
  };

  // x += alpha d, g += alpha h; returns the strip's part of g.g
  auto update = [&](double alpha) -> double {
    double acc[2] = {0.0, 0.0};
    for_each_chunk<K>([&](auto nd, int j0) {
      constexpr int ND = decltype(nd)::value;
      TV<ND> qh, qg, qx;
      tm_ld<ND>(hcol + 2u * j0, qh);
      tm_ld<ND>(gcol + 2u * j0, qg);
      tm_ld<ND>(xcol + 2u * j0, qx);
      tm_wait3<ND>(qh, qg, qx);
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        const double xo = tv_get<ND>(qx, i) + alpha * dl[j0 + i];
        const double go = tv_get<ND>(qg, i) + alpha * tv_get<ND>(qh, i);
        acc[(j0 + i) & 1] += go * (shadow(j0 + i) ? 0.0 : go);
        tm_st1(xcol + 2u * (j0 + i), xo);
        tm_st1(gcol + 2u * (j0 + i), go);
      }
    });
    tm_st_wait();
    return acc[0] + acc[1];
  };

  // d = beta d - g, own strip to shared memory
  auto direction = [&](double beta) {
    for_each_chunk<K>([&](auto nd, int j0) {
      constexpr int ND = decltype(nd)::value;
      TV<ND> qg;
      tm_ld<ND>(gcol + 2u * j0, qg);
      tm_wait<ND>(qg);
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        dl[j0 + i] = beta * dl[j0 + i] - tv_get<ND>(qg, i);
        if (live && !shadow(j0 + i)) dvc[j0 + i] = dl[j0 + i];
      }
    });
  };

  // the block the CG runs on, between global memory (mesh ordering, coalesced) and the padded columns of dv
  auto stage_in = [&](const double* __restrict__ src) {
    for (int c = warp; c < mp; c += nw)
      for (int i = lane; i < mp; i += 32) dv[c * P + i] = src[static_cast<size_t>(c + off) * m + (i + off)];
  };

  // The queue is read one system ahead: the right-hand side and the start vector of the NEXT system are pulled
  // into L2 by bulk prefetches while this one is being solved.
  if (tid == 0) s_next = atomicAdd(a.work_counter, 1);
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_k = s_next;
      if (s_next < a.N) {
        const int nx = atomicAdd(a.work_counter, 1);
        s_next = nx;
        if (nx < a.N) {
          const uint32_t row_bytes = static_cast<uint32_t>(n) * 8u & ~15u;
          l2_prefetch_bulk(a.b + static_cast<size_t>(nx) * n, row_bytes);
          l2_prefetch_bulk(a.x + static_cast<size_t>(nx) * n, row_bytes);
        }
      }
    }
    __syncthreads();
    const int k = s_k;
    if (k >= a.N) break;
    const double* bk = a.b + static_cast<size_t>(k) * n;
    double* xk = a.x + static_cast<size_t>(k) * n;
    if (tid < STENCIL_TABLE)
      tab[tid] = __ldg(a.stencil_ws + (a.diag_stride != 0 ? static_cast<size_t>(k) * STENCIL_TABLE : 0) + tid);
    stage_in(xk);
    __syncthreads();
    c0 = live ? tab[(cx * 3 + 1) * 9 + 0] : 0.0;
    c1 = live ? tab[(cx * 3 + 1) * 9 + 1] : 0.0;
#pragma unroll
    for (int e = 0; e < 3; ++e) cs[e] = tab[(1 * 3 + 1) * 9 + 3 + e];
#pragma unroll
    for (int j = 0; j < K; ++j) dl[j] = dvc[j];  // idle threads read the zero pad
#pragma unroll
    for (int j = 0; j < K; ++j) tm_st1(xcol + 2u * j, dl[j]);
    // ---- g = A x - b
    spmv();
    __syncthreads();  // everybody is done with x in dv
    stage_in(bk);
    __syncthreads();
    double acc0[2] = {0.0, 0.0};
    for_each_chunk<K>([&](auto nd, int j0) {
      constexpr int ND = decltype(nd)::value;
      TV<ND> qh;
      tm_ld<ND>(hcol + 2u * j0, qh);
      tm_wait<ND>(qh);
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        const double go = tv_get<ND>(qh, i) - dvc[j0 + i];
        acc0[(j0 + i) & 1] += go * (shadow(j0 + i) ? 0.0 : go);
        tm_st1(gcol + 2u * (j0 + i), go);
      }
    });
    tm_st_wait();
    double res = sqrt(block_sum12(acc0[0] + acc0[1], red));
    int it = 0;
    bool ok = res <= a.tol;
    if (!ok) {
      // the barrier inside block_sum12 separates the reads of b from this overwrite of dv
      direction(0.0);  // d = -g (0 * x - g)
      double gh = res * res;
      double rgh = FAST ? 1.0 / gh : 0.0;
      for (;;) {
        ++it;
        __syncthreads();
        const double dh = block_sum12(spmv(), red + ST_RED);
        const double alpha = gh / dh;
        const double s2 = block_sum12(update(alpha), red + 2 * ST_RED);
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
        const double beta = FAST ? gh_new * rgh : gh_new / gh;
        gh = gh_new;
        if (FAST) rgh = 1.0 / gh;
        direction(beta);
      }
      if (FAST) res = sqrt(res);
    }
    // ---- write the solution back: tensor memory -> padded columns -> mesh ordering (coalesced)
    __syncthreads();
    for_each_chunk<K>([&](auto nd, int j0) {
      constexpr int ND = decltype(nd)::value;
      TV<ND> qx;
      tm_ld<ND>(xcol + 2u * j0, qx);
      tm_wait<ND>(qx);
#pragma unroll
      for (int i = 0; i < ND; ++i)
        if (live && !shadow(j0 + i)) dvc[j0 + i] = tv_get<ND>(qx, i);
    });
    __syncthreads();
    for (int c = warp; c < mp; c += nw)
      for (int i = lane; i < mp; i += 32) xk[static_cast<size_t>(c + off) * m + (i + off)] = dv[c * P + i];
    if (tid == 0) {
      if (a.iters) a.iters[k] = it;
      if (a.res) a.res[k] = res;
      if (!ok) atomicExch(a.fail_flag, 1);
    }
    // only nodes of the block were written: the pads of dv are still zero for the next system
  }
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem), "r"(static_cast<uint32_t>(p.alloc)));
}

// ------------------------------------------------------------------------------------------------
// Class tables.  Entry e of the row of node (ixp, iyp) of the block [off, off + mp)^2, read from the stored
// diagonals (offsets 0, +1, +m-1, +m, +m+1); couplings that leave the block are zero by definition.
__device__ __forceinline__ double stencil_entry(const double* __restrict__ dg, int ld, int m, int off, int mp, int ixp,
                                                int iyp, int e) {
  const int r = (ixp + off) * m + (iyp + off);
  const bool up = iyp < mp - 1, dn = iyp > 0, rt = ixp < mp - 1, lf = ixp > 0;
  const size_t L = static_cast<size_t>(ld);
  switch (e) {
    case 0: return dg[r];
    case 1: return up ? dg[L + r] : 0.0;
    case 2: return dn ? dg[L + r - 1] : 0.0;
    case 3: return (rt && dn) ? dg[2 * L + r] : 0.0;
    case 4: return rt ? dg[3 * L + r] : 0.0;
    case 5: return (rt && up) ? dg[4 * L + r] : 0.0;
    case 6: return (lf && up) ? dg[2 * L + r - (m - 1)] : 0.0;
    case 7: return lf ? dg[3 * L + r - m] : 0.0;
    default: return (lf && dn) ? dg[4 * L + r - (m + 1)] : 0.0;
  }
}

__device__ __forceinline__ int axis_class(int i, int mp) { return i == 0 ? 0 : (i == mp - 1 ? 2 : 1); }

// One CTA per system (grid-stride).  Threads 0..80 write the table from nine representative nodes; with
// verify != 0 all threads then compare the nine couplings of EVERY node of the block with the table of its
// class, and - Dirichlet blocks - check that the eliminated rows and columns really carry no coupling.
__global__ void __launch_bounds__(128) k_stencil_compress(MeshDims md, int nsys, const double* __restrict__ diag,
                                                           long long diag_stride, int off, int verify,
                                                           double* __restrict__ table, int* __restrict__ bad_flag) {
  __shared__ double tab[STENCIL_TABLE];
  const int m = md.m, mp = m - 2 * off, tid = threadIdx.x;
  for (int k = blockIdx.x; k < nsys; k += gridDim.x) {
    const double* dg = diag + static_cast<size_t>(k) * diag_stride;
    __syncthreads();
    if (tid < STENCIL_TABLE) {
      const int cls = tid / 9, e = tid - 9 * cls, cx = cls / 3, cy = cls - 3 * cx;
      const int ixp = cx == 0 ? 0 : (cx == 1 ? 1 : mp - 1), iyp = cy == 0 ? 0 : (cy == 1 ? 1 : mp - 1);
      const double v = stencil_entry(dg, md.ld, m, off, mp, ixp, iyp, e);
      tab[tid] = v;
      table[static_cast<size_t>(k) * STENCIL_TABLE + tid] = v;
    }
    __syncthreads();
    bool bad = false;
    if (tid < 3) {
      // what row5 of the CG kernel relies on (cheap, so checked on every call): in the inner-row class of every
      // column class up and down couplings coincide, and the couplings to a neighbouring column that exists are
      // those of the inner class, whose right / left pairs coincide
      const double* in = tab + (1 * 3 + 1) * 9;
      const double* c = tab + (tid * 3 + 1) * 9;
      bad |= !(c[1] == c[2]);
      for (int e = 0; e < 3; ++e) {
        bad |= !(in[3 + e] == in[6 + e]);
        if (tid < 2) bad |= !(c[3 + e] == in[3 + e]);
        if (tid > 0) bad |= !(c[6 + e] == in[6 + e]);
      }
    }
    if (verify) {
    for (int rp = tid; rp < mp * mp; rp += 128) {
      const int ixp = rp / mp, iyp = rp - ixp * mp;
      const double* row = tab + (axis_class(ixp, mp) * 3 + axis_class(iyp, mp)) * 9;
#pragma unroll
      for (int e = 0; e < 9; ++e) bad |= !(stencil_entry(dg, md.ld, m, off, mp, ixp, iyp, e) == row[e]);
    }
    if (off) {
      const int ddx[4] = {0, 1, 1, 1}, ddy[4] = {1, -1, 0, 1};
      for (int r = tid; r < md.n; r += 128) {
        const int ix = r / m, iy = r - ix * m;
        const bool me = ix == 0 || iy == 0 || ix == m - 1 || iy == m - 1;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int jx = ix + ddx[d], jy = iy + ddy[d];
          if (jx < 0 || jy < 0 || jx >= m || jy >= m) continue;
          const bool other = jx == 0 || jy == 0 || jx == m - 1 || jy == m - 1;
          if (me || other) bad |= !(dg[static_cast<size_t>(d + 1) * md.ld + r] == 0.0);
        }
      }
    }
    }
    if (bad) atomicExch(bad_flag, 1);
  }
}

template <int K>
int launch_plan(cudaStream_t s, const CgArgs& a, const StencilPlan& p, int num_sms) {
  const size_t smem = plan_smem_bytes(p);
  auto kern = k_cg_stencil<K>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  // CTAs per SM: tensor-memory columns (not known to the occupancy calculator), registers, shared memory
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int per_sm = 512 / p.alloc;
  const int by_regs = 65536 / (fa.numRegs * p.block), by_smem = static_cast<int>((227u * 1024u) / (smem + fa.sharedSizeBytes + 1024));
  if (per_sm > by_regs) per_sm = by_regs;
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > 2048 / p.block) per_sm = 2048 / p.block;
  if (const char* e = std::getenv("MSGPU_STENCIL_CTAS")) {  // A/B: fewer CTAs per SM than fit
    const int want = std::atoi(e);
    if (want >= 1 && want < per_sm) per_sm = want;
  }
  if (per_sm < 1) per_sm = 1;
  // shared-memory carve-out that holds per_sm CTAs (the rest stays L1)
  int carve = static_cast<int>((per_sm * (smem + fa.sharedSizeBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
  if (carve > 100) carve = 100;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  if (std::getenv("MSGPU_DEBUG_LAUNCH"))
    std::fprintf(stderr,
                 "[msgpu] k_cg_stencil<%d>: block %d (S %d, mp %d, pitch %d, overlap %d), %d regs, smem %zu B, TMEM %d cols -> %d "
                 "CTAs per SM\n",
                 K, p.block, p.S, p.mp, p.P, p.ov, fa.numRegs, smem, p.alloc, per_sm);
  long long grid = static_cast<long long>(num_sms) * per_sm;
  if (grid > a.N) grid = a.N;
  kern<<<static_cast<int>(grid), p.block, smem, s>>>(a, p);
  return 1;
}

}  // namespace

bool stencil_supported(const MeshDims& md, int dirichlet) {
  StencilPlan p;
  return make_plan(md, dirichlet, p);
}

int launch_stencil_compress(cudaStream_t s, const CgArgs& a, int verify) {
  const int nsys = a.diag_stride != 0 ? a.N : 1;
  int grid = nsys < 148 * 8 ? nsys : 148 * 8;
  k_stencil_compress<<<grid, 128, 0, s>>>(a.md, nsys, a.diag, a.diag_stride, a.dirichlet ? 1 : 0, verify, a.stencil_ws,
                                          a.stencil_bad);
  return 1;
}

int launch_cg_stencil(cudaStream_t s, const CgArgs& a, int num_sms) {
  StencilPlan p;
  if (!make_plan(a.md, a.dirichlet, p)) return -1;
  switch (p.K) {
    case 9: return launch_plan<9>(s, a, p, num_sms);
    case 10: return launch_plan<10>(s, a, p, num_sms);
    case 11: return launch_plan<11>(s, a, p, num_sms);
    case 12: return launch_plan<12>(s, a, p, num_sms);
    case 13: return launch_plan<13>(s, a, p, num_sms);
  }
  return -1;
}

}  // namespace msgpu
