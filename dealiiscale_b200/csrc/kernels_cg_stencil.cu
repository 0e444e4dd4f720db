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
// With the matrix gone there is nothing left to stream per CG step but the neighbours of the search direction, and
// the SpMV - 1400 of the 3030 cycles of a k_cg_strip_tmem step, bound by the tensor-memory reads - becomes FP64 work
// on registers, provided the shared-memory traffic of the neighbour exchange stays below it (one wavefront per cycle
// and SM; 13 x 1 strips needed 1543 wavefronts per step, see profiles/r2_a_stencil_tmem_variant.md):
//   * a thread owns a TILE of W = 3 adjacent mesh columns x KR = 4 consecutive nodes: 18 neighbour loads and 12 stores
//     per 12 nodes.  Tile rows: S strips per column, the last strip anchored at the end of the column, so the boundary
//     nodes of a column sit at compile-time positions (row 0 of the first strip, row KR-1 of the last) and the
//     S*KR - mp rows the last strip shares with its neighbour are computed twice, bit-identically, and counted once.
//     Tile columns past the mesh (mp not a multiple of 3) compute zeros;
//   * x, g, h = A d and the own tile of d are registers; d is also in shared memory for the neighbours (column pitch
//     P with a zero pad that doubles as the missing neighbours of boundary nodes, so there are no index tests);
//   * every coupling is a register: k_stencil_compress checks that the class table has the form
//         self(cx, cy),  vertical(cx),  horizontal(cy),  two diagonal values
//     (what a symmetric operator with per-system constant coefficients and boundary terms along the faces gives), so a
//     thread needs 9 + 3 + 2 values of its own and 3 that are the same for everybody;
//   * one CTA of <= 384 threads per SM for 65 x 65 nodes, more for smaller meshes.
// Thread t -> (strip s = t / TC, tile column tc = t % TC): adjacent lanes own adjacent tiles, shared-memory accesses
// of a warp are W P doubles apart (W, P odd: conflict free).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "cg_common.cuh"
#include "checked.cuh"
#include "kernels.h"

namespace msgpu {

namespace {

using cgc::l2_prefetch_bulk;
using cgc::warp_sum;

#ifdef MSGPU_CHECKED
__device__ int g_checked_selftest = 0;  // MSGPU_CHECKED_SELFTEST=1: one deliberate out-of-window read (tests)
#endif

constexpr int ST_MAX_THREADS = 256;  // 8 warps: 255 registers (one CTA per SM) or 128 (two, vectors in tensor memory)
constexpr int ST_RED = 8;            // slots of one reduction buffer (= warps per CTA)

struct StencilPlan {
  int mp;     // nodes per axis of the block the CG runs on: m, or m - 2 when Dirichlet rows are left out
  int off;    // first mesh line of that block (0 | 1)
  int W, KR;  // tile of a thread: columns x rows
  int TC;     // tiles per mesh row: ceil(mp / W)
  int S;      // strips per mesh column: ceil(mp / KR)
  int P;      // pitch of a column of d in shared memory (odd, >= mp + 1)
  int ov;     // rows the last strip of a column shares with the strip before it (S*KR - mp)
  int block;  // CTA size
  int rows_fastest;  // thread -> tile: adjacent lanes own tiles adjacent in the column (1) or in the row (0) - whichever
                     // makes the distance of their shared-memory accesses an odd number of doubles (conflict free)
  int alloc;  // tensor-memory variant: columns a CTA allocates (power of two) ...
  int cw;     // ... and columns per warp (x, g, h of a tile: 6 * W * KR)
};

// Tile shapes, best first: no rows shared between strips and no tile columns past the mesh (63 = 21 x 3 = 9 x 7 nodes:
// the 64-cell Dirichlet block), then anything.  At most 21 nodes per tile: 6 x 21 tensor-memory columns for each of
// the two warps of a lane quadrant, two CTAs per SM.  (Measured and dropped: 4 x 5 tiles for 65 = 13 x 5 nodes - no
// shared rows, but 42 coupling registers next to 40 for d leave ptxas too little room at 128: 9.7 against 8.9 ms.)
struct TileShape {
  int W, KR, any_overlap, any_dead_columns;
};
constexpr int ST_NSHAPES = 2;
constexpr TileShape ST_SHAPES[ST_NSHAPES] = {{3, 7, 0, 0}, {3, 6, 1, 1}};

bool make_plan(const MeshDims& md, int dirichlet, StencilPlan& p, int* shape) {
  const int off = dirichlet ? 1 : 0, mp = md.m - 2 * off;
  if (mp < 48) return false;  // smaller systems: several CTAs of the plain strip kernel per SM do as well or better
  for (int i = 0; i < ST_NSHAPES; ++i) {
    const TileShape& t = ST_SHAPES[i];
    const int TC = (mp + t.W - 1) / t.W, S = (mp + t.KR - 1) / t.KR;
    if (TC * S > ST_MAX_THREADS) continue;
    if (!t.any_overlap && S * t.KR != mp) continue;
    if (!t.any_dead_columns && TC * t.W != mp) continue;
    p.mp = mp;
    p.off = off;
    p.W = t.W;
    p.KR = t.KR;
    p.TC = TC;
    p.S = S;
    p.ov = S * t.KR - mp;
    // Adjacent lanes: tiles W*P doubles apart (next tile of the strip) or KR doubles apart (next strip of the tile
    // column); 64-bit accesses of a half-warp are conflict free when that distance is odd.  The column pitch P is
    // odd and, where possible, such that the tile after a seam continues the progression (mod 16 doubles).
    p.rows_fastest = (t.W % 2 == 0) ? 1 : 0;
    p.P = (mp + 1) | 1;
    for (int P = p.P; P < p.P + 32; P += 2) {
      const int jump = p.rows_fastest ? (t.W * P - S * t.KR) - t.KR : (t.W * P * TC - t.KR);
      if (((jump % 16) + 16) % 16 == 0) {
        p.P = P;
        break;
      }
    }
    p.block = (TC * S + 31) & ~31;
    p.cw = 6 * t.W * t.KR;
    p.alloc = 32;
    while (p.alloc < (p.block / 32 + 3) / 4 * p.cw) p.alloc *= 2;
    if (shape) *shape = i;
    return true;
  }
  return false;
}

// dv[ix * P + iy] for ix in [-1, mp + 2 W]: one guard column in front, tile columns past the mesh and the tile of idle
// threads behind; 2 doubles in front of everything for the (-1, -1) corner of the first tile
size_t plan_dv_len(const StencilPlan& p) {
  return (static_cast<size_t>(2) + static_cast<size_t>(p.mp + 2 * p.W + 2) * p.P + 1) & ~static_cast<size_t>(1);
}
size_t plan_smem_bytes(const StencilPlan& p) { return (plan_dv_len(p) + 3 * ST_RED + STENCIL_TABLE + 1) * sizeof(double); }

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
__device__ __forceinline__ void tm_wait2(TV<ND>& a, TV<ND>& b) {
  if constexpr (ND == 4) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(a.r[4]), "+r"(a.r[5]), "+r"(a.r[6]),
                   "+r"(a.r[7]), "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]),
                   "+r"(b.r[6]), "+r"(b.r[7]));
  } else if constexpr (ND == 2) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]),
                   "+r"(b.r[3]));
  } else {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(b.r[0]), "+r"(b.r[1]));
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

// f(integral_constant<int, ND>, e0) for the chunks [e0, e0 + ND) of E values: fours, then a pair, then one
template <int E, class F>
__device__ __forceinline__ void for_each_chunk(F&& f) {
#pragma unroll
  for (int c = 0; c < E / 4; ++c) f(std::integral_constant<int, 4>{}, 4 * c);
  if constexpr ((E % 4) >= 2) f(std::integral_constant<int, 2>{}, E - (E % 4));
  if constexpr ((E % 2) == 1) f(std::integral_constant<int, 1>{}, E - 1);
}

// Sum over the CTA, result in every thread: warp partials into `red` (ST_RED slots; the slots of warps the CTA does
// not have stay zero), one barrier, then every thread adds the eight in one fixed tree (bit-reproducible; a tree of
// additions has a shorter dependency chain than a second round of shuffles).
// `red` must not be written again before the next barrier that follows this call (callers rotate three buffers).
__device__ __forceinline__ double block_sum8(double v, double* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t[8];
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const double2 q = *reinterpret_cast<const double2*>(red + i);
    t[i] = q.x;
    t[i + 1] = q.y;
  }
  return ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
}

// One matrix row: nine products, one chain, fixed order: self, up, down, right-down, left-up, right, left, right-up,
// left-down (a thread has 12 independent rows in flight; joining two chains would cost one more FP64 addition per row).
__device__ __forceinline__ double stencil_row(double cself, double cv, double ch, double cda, double cdb, double self,
                                              double above, double below, double l_dn, double l_md, double l_up,
                                              double r_dn, double r_md, double r_up) {
  double s = cself * self;
  s += cv * above;
  s += cv * below;
  s += cda * r_dn;
  s += cda * l_up;
  s += ch * r_md;
  s += ch * l_md;
  s += cdb * r_up;
  s += cdb * l_dn;
  return s;
}

// TM = false: x, g, h in registers, one CTA of 256 threads per SM.
// TM = true:  x, g, h in TENSOR MEMORY (6 W KR columns per thread, streamed through registers four values at a time:
//             tcgen05.ld / tcgen05.st, SASS LDTM / STTM), 128 registers, TWO CTAs per SM: while one system sits in
//             its reductions and scalar chains the other one keeps the FP64 pipe busy.
// W x KR: tile of a thread.  OVN / DCN: rows shared between the last two strips of a column / tile columns past the
// mesh - an exact count (their masks then touch only those rows / columns; 0: none, compiled out) or -1: whatever the
// plan says (masks on every row / column that could be affected: ~110 selects per CG step).
// (Measured and dropped: the couplings of the first and last row of a tile read from the class table in shared memory
// at their place of use instead of sitting in 16 registers: 8.63 against 8.28 ms; all of g loaded from tensor memory
// at once in the direction update instead of four values at a time: 8.37 against 8.28 ms.)
// XL (tensor-memory variant): x += alpha d is done in the direction phase behind the second reduction - where the
// loads of g leave the FP64 pipe idle anyway - instead of in the update phase between the two reductions of a step
// (8.27 -> 8.18 ms on the bench workload, the same bits: the operations and their operands do not change).  A system
// that leaves the loop right after its last update adds the pending alpha d when it writes x back.
// (Measured and dropped on top of it: the first four values of g fetched while the block sum of d.h is formed and kept
// in registers through the second reduction, the first four of x fetched during the latter: 8.36 ms.)
// (Requesting the first values of h and g before alpha = gh / d.h is formed changes nothing: 8.19 ms; g and x written
// back two doubles per tcgen05.st: 8.37 ms, the pairs cost register moves again.)
template <bool TM, int W, int KR, int OVN, int DCN, bool XL = TM>
__global__ void __launch_bounds__(ST_MAX_THREADS, TM ? 2 : 1)
    k_cg_stencil(const __grid_constant__ CgArgs a, const __grid_constant__ StencilPlan p) {
  if (a.gate != nullptr && ((*a.gate == 0) != (a.gate_run_if_zero != 0))) return;  // the other kernel does the work
  constexpr bool FAST = MSGPU_FAST_SCALARS;
  constexpr int E = W * KR;
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_k, s_next;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int mp = p.mp, P = p.P, m = a.md.m, n = a.md.n, off = p.off, ov = p.ov;
  const int dvlen = (2 + (mp + 2 * W + 2) * P + 1) & ~1;
  const CheckedPtr<double> dv = checked(sm + 2 + P, sm, sm + dvlen);  // everything nobody writes stays zero
  double* const red = sm + dvlen;  // 3 x ST_RED, 16-byte aligned
  double* const tab = red + 3 * ST_RED;

  const int s = p.rows_fastest ? tid % p.S : tid / p.TC;
  const int tcol = p.rows_fastest ? tid / p.S : tid % p.TC;
  const bool live = tid < p.S * p.TC;  // idle threads sit on columns mp + W - 1 ...: zeros there and on both sides
  const int ix0 = live ? tcol * W : mp + W - 1;
  const bool first = live && s == 0, last = live && s == p.S - 1;
  const CheckedPtr<double> dvt = dv + (ix0 * P + (live ? (last ? mp - KR : s * KR) : 0));  // node (0, 0) of the tile
#ifdef MSGPU_CHECKED
  if (g_checked_selftest && tid == 0 && dv[-(2 + P) - 1] == 1.2345e-300) return;  // must trip the window check
#endif
  bool col_live[W];  // tile columns past the mesh compute zeros and write nothing
#pragma unroll
  for (int c = 0; c < W; ++c) col_live[c] = live && (DCN == 0 || ix0 + c < mp);
  // can tile column c lie past the mesh at all
  auto maybe_dead = [](int c) -> bool { return DCN < 0 ? c > 0 : c >= W - DCN; };
  // rows [0, ov) of a last strip belong to the strip before it: same arithmetic there and here, counted there
  auto shadow = [&](int j) -> bool {
    if (OVN == 0) return false;
    if (OVN > 0) return j < OVN && last;
    return j < KR - 1 && last && j < ov;
  };
  (void)ov;

  if (TM && warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(&s_tmem))),
                 "r"(static_cast<uint32_t>(p.alloc)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  for (int i = tid; i < dvlen + 3 * ST_RED + STENCIL_TABLE + 1; i += blockDim.x) sm[i] = 0.0;
  uint32_t xcol = 0, gcol = 0, hcol = 0;
  if constexpr (TM) {
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    // lane quadrant of this warp in bits 31:16, column block of this warp in bits 15:0
    xcol = s_tmem + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) +
           static_cast<uint32_t>(warp >> 2) * static_cast<uint32_t>(p.cw);
    gcol = xcol + 2u * E;
    hcol = xcol + 4u * E;
    MSGPU_CHECK(6 * E <= p.cw && (warp >> 2) * p.cw + p.cw <= p.alloc);
  }
  // checked builds: every tensor-memory access must stay inside the 6 E columns of this warp
  [[maybe_unused]] const uint32_t tm_lo = xcol & 0xffffu, tm_hi = tm_lo + 6u * E;
#define MSGPU_TM(a, nd) MSGPU_CHECK_TMEM(a, 2u * (nd), tm_lo, tm_hi)

  double xx[W][KR], gg[W][KR], dl[W][KR], hh[W][KR];  // solution, residual, search direction, A d: this thread's tile
                                                      // (TM: only dl; the others live in tensor memory)
  double cself[W][3];  // diagonal entries of the tile's columns: row 0, inner rows, row KR-1 of the tile
  double cvert[W];     // up / down coupling of the tile's columns
  double chor[3];      // left / right coupling for row 0, inner rows, row KR-1 of the tile
  double cda, cdb;     // diagonal couplings: right-down | left-up, right-up | left-down (the same for every thread)

  // the block the CG runs on, between global memory (mesh ordering, coalesced) and the padded columns of dv
  auto stage_in = [&](const double* __restrict__ src) {
    for (int c = warp; c < mp; c += nw)
      for (int i = lane; i < mp; i += 32) dv[c * P + i] = src[static_cast<size_t>(c + off) * m + (i + off)];
  };
// every node (c, j) of the tile, fully unrolled (the loops sit at their place of use: a lambda called from several
// places does not always get unrolled, and a loop that is not turns the register tiles into local memory)
#define MSGPU_TILE(c, j)                                \
  _Pragma("unroll") for (int e_ = 0; e_ < W * KR; ++e_) \
    if (const int c = e_ / KR; true)                    \
      if (const int j = e_ % KR; true)
// own tile of x to shared memory; rows shared with the strip before: the owner writes.  Of d only the nodes on the
// rim of the tile: nobody else reads the inner ones
#define MSGPU_PUBLISH(v) MSGPU_TILE(c, j) if (col_live[c] && !shadow(j)) dvt[c * P + j] = v[c][j]
// (with strips that share rows the row below / above a tile can be an inner row of the neighbour: publish everything)
#define MSGPU_RIM(c, j) (OVN != 0 || (c) == 0 || (c) == W - 1 || (j) == 0 || (j) == KR - 1)
#define MSGPU_PUBLISH_RIM(v) \
  MSGPU_TILE(c, j) if (MSGPU_RIM(c, j) && col_live[c] && !shadow(j)) dvt[c * P + j] = v[c][j]

  // The queue is read one system ahead: the right-hand side and the start vector of the NEXT system are pulled
  // into L2 by bulk prefetches while this one is being solved.
  if (tid == 0) s_next = atomicAdd(a.work_counter, 1);
#ifdef MSGPU_STENCIL_STAMPS
  long long st_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0}, st_its = 0;  // debug build: clock64 cycles per phase
#endif
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
          l2_prefetch_bulk((a.x_in ? a.x_in : a.x) + static_cast<size_t>(nx) * n, row_bytes);
        }
      }
    }
    __syncthreads();
    const int k = s_k;
    if (k >= a.N) break;
    const double* bk = a.b + static_cast<size_t>(k) * n;
    double* xk = a.x + static_cast<size_t>(k) * n;
    const double* xin = (a.x_in ? a.x_in : a.x) + static_cast<size_t>(k) * n;  // start vector (msgpu_solve, bio)
    if (tid < STENCIL_TABLE)
      tab[tid] = __ldg(a.stencil_ws + (a.diag_stride != 0 ? static_cast<size_t>(k) * STENCIL_TABLE : 0) + tid);
    stage_in(xin);
    __syncthreads();
    // couplings of this tile from the class table T[cx][cy][e] (form checked by k_stencil_compress)
    {
      const int cy_lo = first ? 0 : 1, cy_hi = last ? 2 : 1;
#pragma unroll
      for (int c = 0; c < W; ++c) {
        const int ix = ix0 + c, cx = ix == 0 ? 0 : (ix == mp - 1 ? 2 : 1);
        cself[c][1] = col_live[c] ? tab[(cx * 3 + 1) * 9] : 0.0;
        cvert[c] = col_live[c] ? tab[(cx * 3 + 1) * 9 + 1] : 0.0;
        cself[c][0] = col_live[c] ? tab[(cx * 3 + cy_lo) * 9] : 0.0;
        cself[c][2] = col_live[c] ? tab[(cx * 3 + cy_hi) * 9] : 0.0;
      }
      chor[0] = tab[(1 * 3 + cy_lo) * 9 + 4];
      chor[1] = tab[(1 * 3 + 1) * 9 + 4];
      chor[2] = tab[(1 * 3 + cy_hi) * 9 + 4];
      cda = tab[(1 * 3 + 1) * 9 + 3];
      cdb = tab[(1 * 3 + 1) * 9 + 5];
    }
    MSGPU_TILE(c, j) {
      dl[c][j] = dvt[c * P + j];  // idle threads and columns past the mesh read the zero pad
      if constexpr (TM) {
        MSGPU_TM(xcol + 2u * (c * KR + j), 1);
        tm_st1(xcol + 2u * (c * KR + j), dl[c][j]);
      } else {
        xx[c][j] = dl[c][j];
        gg[c][j] = 0.0;
      }
    }
    // The first pass through the loop computes g = A x - b (d = x is in shared memory already), the following
    // ones are CG steps: one place of use for the SpMV.
    bool start = true, ok = false;
    int it = 0;
    double res = 0.0, gh = 0.0, rgh = 0.0, alpha = 0.0;
    [[maybe_unused]] bool x_pending = false;  // XL: the x update of the step just taken has not been done yet
#ifdef MSGPU_STENCIL_STAMPS
    long long st_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, st_c = clock64();
#define MSGPU_STAMP(i)                                    \
  {                                                       \
    const long long now_ = clock64();                     \
    st_t[i] += now_ - st_c;                               \
    st_c = now_;                                          \
  }
#else
#define MSGPU_STAMP(i)
#endif
    for (;;) {
      __syncthreads();
      MSGPU_STAMP(0)  // closing barrier of the previous step
      // ---- hh = A dl for this tile, and the tile's part of d.h
      double dh_part;
      {
        double L[KR + 2], R[KR + 2], B[W], A[W];
#pragma unroll
        for (int j = 0; j < KR + 2; ++j) {
          L[j] = dvt[-P + j - 1];
          R[j] = dvt[W * P + j - 1];
        }
#pragma unroll
        for (int c = 0; c < W; ++c) {
          B[c] = dvt[c * P - 1];
          A[c] = dvt[c * P + KR];
        }
        // value of d at tile position (c, j), c in [-1, W], j in [-1, KR]
        auto D = [&](int c, int j) -> double {
          if (c < 0) return L[j + 1];
          if (c >= W) return R[j + 1];
          if (j < 0) return B[c];
          if (j >= KR) return A[c];
          return dl[c][j];
        };
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        MSGPU_TILE(c, j) {
          const int rc = j == 0 ? 0 : (j == KR - 1 ? 2 : 1);
          double h = stencil_row(cself[c][rc], cvert[c], chor[rc], cda, cdb, D(c, j), D(c, j + 1), D(c, j - 1),
                                 D(c - 1, j - 1), D(c - 1, j), D(c - 1, j + 1), D(c + 1, j - 1), D(c + 1, j),
                                 D(c + 1, j + 1));
          if (maybe_dead(c)) h = col_live[c] ? h : 0.0;
          if constexpr (TM) {
            MSGPU_TM(hcol + 2u * (c * KR + j), 1);
            tm_st1(hcol + 2u * (c * KR + j), h);
          } else {
            hh[c][j] = h;
          }
          acc[(c * KR + j) & 3] += dl[c][j] * (shadow(j) ? 0.0 : h);
        }
        dh_part = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      }
#ifdef MSGPU_STENCIL_STAMPS
      if (dh_part == 1.2345e-300) st_t[7] += 1;  // the stamp below must wait for the SpMV
#endif
      MSGPU_STAMP(1)  // SpMV + own part of d.h
      if (start) {
        start = false;
        __syncthreads();  // everybody is done with x in dv
        stage_in(bk);
        __syncthreads();
        double part = 0.0;
        if constexpr (TM) {
          tm_st_wait();
          for_each_chunk<E>([&](auto nd, int e0) {
            constexpr int ND = decltype(nd)::value;
            TV<ND> qh;
            MSGPU_TM(hcol + 2u * e0, ND);
            tm_ld<ND>(hcol + 2u * e0, qh);
            tm_wait<ND>(qh);
#pragma unroll
            for (int i = 0; i < ND; ++i) {
              const int c = (e0 + i) / KR, j = (e0 + i) % KR;
              const double go = tv_get<ND>(qh, i) - dvt[c * P + j];
              part += go * (shadow(j) ? 0.0 : go);
              MSGPU_TM(gcol + 2u * (e0 + i), 1);
              tm_st1(gcol + 2u * (e0 + i), go);
              dl[c][j] = -go;
            }
          });
        } else {
          MSGPU_TILE(c, j) {
            gg[c][j] = hh[c][j] - dvt[c * P + j];
            part += gg[c][j] * (shadow(j) ? 0.0 : gg[c][j]);
            dl[c][j] = -gg[c][j];
          }
        }
        res = sqrt(block_sum8(part, red));
        if (res <= a.tol) {
          ok = true;
          break;
        }
        // the barrier inside block_sum8 separates the reads of b from this overwrite of dv (all of it: the inner
        // nodes of the tiles still hold b)
        MSGPU_PUBLISH(dl);
        gh = res * res;
        rgh = FAST ? 1.0 / gh : 0.0;
        if (FAST) res = gh;
        continue;
      }
      ++it;
      const double dh = block_sum8(dh_part, red + ST_RED);
#ifdef MSGPU_STENCIL_STAMPS
      if (dh == 1.2345e-300) st_t[7] += 1;
#endif
      MSGPU_STAMP(2)  // first reduction
      alpha = gh / dh;
#ifdef MSGPU_STENCIL_STAMPS
      if (alpha == 1.2345e-300) st_t[7] += 1;
#endif
      MSGPU_STAMP(3)  // alpha
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      if constexpr (TM) {
        auto update = [&](auto nd, int e0, auto& qh, auto& qg, auto& qx) {
          constexpr int ND = decltype(nd)::value;
#pragma unroll
          for (int i = 0; i < ND; ++i) {
            const int c = (e0 + i) / KR, j = (e0 + i) % KR;
            const double go = tv_get<ND>(qg, i) + alpha * tv_get<ND>(qh, i);
            acc[(e0 + i) & 3] += go * (shadow(j) ? 0.0 : go);
            MSGPU_TM(gcol + 2u * (e0 + i), 1);
            tm_st1(gcol + 2u * (e0 + i), go);
            if constexpr (!XL) {
              const double xo = tv_get<ND>(qx, i) + alpha * dl[c][j];
              MSGPU_TM(xcol + 2u * (e0 + i), 1);
              tm_st1(xcol + 2u * (e0 + i), xo);
            }
          }
        };
        auto chunk = [&](auto nd, int e0) {
          constexpr int ND = decltype(nd)::value;
          TV<ND> qh, qg, qx;
          MSGPU_TM(hcol + 2u * e0, ND);
          tm_ld<ND>(hcol + 2u * e0, qh);
          MSGPU_TM(gcol + 2u * e0, ND);
          tm_ld<ND>(gcol + 2u * e0, qg);
          if constexpr (!XL) {
            MSGPU_TM(xcol + 2u * e0, ND);
            tm_ld<ND>(xcol + 2u * e0, qx);
            tm_wait3<ND>(qh, qg, qx);
          } else {
            tm_wait2<ND>(qh, qg);
          }
          update(nd, e0, qh, qg, qx);
        };
        tm_st_wait();  // h of this step (and x, g of the last one) have landed
        for_each_chunk<E>(chunk);
        if constexpr (XL) x_pending = true;
      } else {
        MSGPU_TILE(c, j) {
          xx[c][j] += alpha * dl[c][j];
          gg[c][j] += alpha * hh[c][j];
          acc[(c * KR + j) & 3] += gg[c][j] * (shadow(j) ? 0.0 : gg[c][j]);
        }
      }
      const double part2 = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#ifdef MSGPU_STENCIL_STAMPS
      if (part2 == 1.2345e-300) st_t[7] += 1;
#endif
      MSGPU_STAMP(4)  // x, g updates + own part of g.g
      const double s2 = block_sum8(part2, red + 2 * ST_RED);
#ifdef MSGPU_STENCIL_STAMPS
      if (s2 == 1.2345e-300) st_t[7] += 1;
#endif
      MSGPU_STAMP(5)  // second reduction
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
      if constexpr (TM) {
        tm_st_wait();
        for_each_chunk<E>([&](auto nd, int e0) {
          constexpr int ND = decltype(nd)::value;
          TV<ND> qg, qx;
          MSGPU_TM(gcol + 2u * e0, ND);
          tm_ld<ND>(gcol + 2u * e0, qg);
          if constexpr (XL) {
            MSGPU_TM(xcol + 2u * e0, ND);
            tm_ld<ND>(xcol + 2u * e0, qx);
            tm_wait2<ND>(qg, qx);
          } else {
            tm_wait<ND>(qg);
          }
#pragma unroll
          for (int i = 0; i < ND; ++i) {
            const int c = (e0 + i) / KR, j = (e0 + i) % KR;
            if constexpr (XL) {  // x += alpha d with the direction of the step just taken, before it is replaced
              const double xo = tv_get<ND>(qx, i) + alpha * dl[c][j];
              MSGPU_TM(xcol + 2u * (e0 + i), 1);
              tm_st1(xcol + 2u * (e0 + i), xo);
            }
            dl[c][j] = beta * dl[c][j] - tv_get<ND>(qg, i);
            if (MSGPU_RIM(c, j) && col_live[c] && !shadow(j)) dvt[c * P + j] = dl[c][j];
          }
        });
        if constexpr (XL) x_pending = false;
      } else {
        MSGPU_TILE(c, j) dl[c][j] = beta * dl[c][j] - gg[c][j];
        MSGPU_PUBLISH_RIM(dl);
      }
      MSGPU_STAMP(6)  // beta, new direction, publish
    }
    if (FAST && it > 0) res = sqrt(res);
#ifdef MSGPU_STENCIL_STAMPS
    st_its += it;
    for (int i = 0; i < 8; ++i) st_sum[i] += st_t[i];
#endif
    // ---- write the solution back: registers -> padded columns -> mesh ordering (coalesced)
    __syncthreads();
    if constexpr (TM) {
      tm_st_wait();
      for_each_chunk<E>([&](auto nd, int e0) {
        constexpr int ND = decltype(nd)::value;
        TV<ND> qx;
        MSGPU_TM(xcol + 2u * e0, ND);
        tm_ld<ND>(xcol + 2u * e0, qx);
        tm_wait<ND>(qx);
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          const int c = (e0 + i) / KR, j = (e0 + i) % KR;
          double xv = tv_get<ND>(qx, i);
          if constexpr (XL) xv = x_pending ? xv + alpha * dl[c][j] : xv;  // the loop was left before this update
          if (col_live[c] && !shadow(j)) dvt[c * P + j] = xv;
        }
      });
    } else {
      MSGPU_PUBLISH(xx);
    }
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
#ifdef MSGPU_STENCIL_STAMPS
  // debug build: CTA 0 leaves its mean cycles per CG step and phase in res[0..6] (instead of the residual norms)
  if (blockIdx.x == 0 && tid == 0 && a.res && st_its > 0)
    for (int i = 0; i < 7; ++i) a.res[i] = static_cast<double>(st_sum[i]) / static_cast<double>(st_its);
#endif
#undef MSGPU_STAMP
#undef MSGPU_PUBLISH_RIM
#undef MSGPU_RIM
#undef MSGPU_PUBLISH
#undef MSGPU_TILE
#undef MSGPU_TM
  if constexpr (TM) {
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem),
                   "r"(static_cast<uint32_t>(p.alloc)));
  }
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
    if (tid < STENCIL_TABLE) {
      // the form the CG kernel relies on (cheap, so checked on every call): vertical couplings depend on the column
      // class only, horizontal ones on the row class only, the diagonal ones on nothing; a neighbour that does not
      // exist has a zero entry
      const int cls = tid / 9, e = tid - 9 * cls, cx = cls / 3, cy = cls - 3 * cx;
      const bool up = cy < 2, dn = cy > 0, rt = cx < 2, lf = cx > 0;
      const double* in = tab + (1 * 3 + 1) * 9;
      double want = tab[tid];
      switch (e) {
        case 1: want = up ? tab[(cx * 3 + 1) * 9 + 1] : 0.0; break;
        case 2: want = dn ? tab[(cx * 3 + 1) * 9 + 1] : 0.0; break;
        case 3: want = (rt && dn) ? in[3] : 0.0; break;
        case 4: want = rt ? tab[(1 * 3 + cy) * 9 + 4] : 0.0; break;
        case 5: want = (rt && up) ? in[5] : 0.0; break;
        case 6: want = (lf && up) ? in[3] : 0.0; break;
        case 7: want = lf ? tab[(1 * 3 + cy) * 9 + 4] : 0.0; break;
        case 8: want = (lf && dn) ? in[5] : 0.0; break;
        default: break;
      }
      bad |= !(tab[tid] == want);
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

template <bool TM, int W, int KR, int OVN, int DCN, bool XL = TM>
int launch_plan(cudaStream_t s, const CgArgs& a, const StencilPlan& p, int num_sms, const CgSwitches& sw) {
  const size_t smem = plan_smem_bytes(p);
  auto kern = k_cg_stencil<TM, W, KR, OVN, DCN, XL>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  // CTAs per SM: registers, shared memory, threads (the kernel is persistent: a CTA that becomes resident late
  // finds the queue drained)
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int per_sm = 65536 / (fa.numRegs * ((p.block + 127) / 128 * 128));  // registers are handed out per 4 warps
  const int by_smem = static_cast<int>((227u * 1024u) / (smem + fa.sharedSizeBytes + 1024));
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > 2048 / p.block) per_sm = 2048 / p.block;
  if (TM && per_sm > 512 / p.alloc) per_sm = 512 / p.alloc;  // tensor-memory columns
  if (sw.stencil_ctas >= 1 && sw.stencil_ctas < per_sm) per_sm = sw.stencil_ctas;  // A/B: fewer CTAs per SM than fit
  if (per_sm < 1) per_sm = 1;
  // shared-memory carve-out that holds per_sm CTAs (the rest stays L1)
  int carve = static_cast<int>((per_sm * (smem + fa.sharedSizeBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
  if (carve > 100) carve = 100;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  if (sw.debug_launch)
    std::fprintf(stderr,
                 "[msgpu] k_cg_stencil<TM %d, OVN %d, DCN %d>: block %d (%d x %d tiles of %d x %d nodes, mp %d, pitch %d, overlap %d), %d regs, smem "
                 "%zu B -> %d CTAs per SM\n",
                 TM ? 1 : 0, OVN, DCN, p.block, p.TC, p.S, W, KR, p.mp, p.P, p.ov, fa.numRegs, smem, per_sm);
  long long grid = static_cast<long long>(num_sms) * per_sm;
  if (grid > a.N) grid = a.N;
  kern<<<static_cast<int>(grid), p.block, smem, s>>>(a, p);
  return 1;
}

}  // namespace

bool stencil_supported(const MeshDims& md, int dirichlet) {
  StencilPlan p;
  return make_plan(md, dirichlet, p, nullptr);
}

int launch_stencil_compress(cudaStream_t s, const CgArgs& a, int verify) {
  const int nsys = a.diag_stride != 0 ? a.N : 1;
  int grid = nsys < 148 * 8 ? nsys : 148 * 8;
  k_stencil_compress<<<grid, 128, 0, s>>>(a.md, nsys, a.diag, a.diag_stride, a.dirichlet ? 1 : 0, verify, a.stencil_ws,
                                          a.stencil_bad);
  return 1;
}

int launch_cg_stencil(cudaStream_t s, const CgArgs& a, int num_sms, const CgSwitches& sw) {
#ifdef MSGPU_CHECKED
  if (std::getenv("MSGPU_CHECKED_SELFTEST") != nullptr) {
    const int one = 1;
    cudaMemcpyToSymbol(g_checked_selftest, &one, sizeof one);
  }
#endif
  StencilPlan p;
  int shape = 0;
  if (!make_plan(a.md, a.dirichlet, p, &shape)) return -1;
  const bool use_tm = sw.stencil_tmem != 0;  // 0: vectors in registers, one system per SM (A/B)
  const int dc = p.TC * p.W - p.mp;
  if (shape == 0)
    return use_tm ? launch_plan<true, 3, 7, 0, 0>(s, a, p, num_sms, sw) : launch_plan<false, 3, 7, 0, 0>(s, a, p, num_sms, sw);
  // 65 x 65 nodes (64 x 64 cells, Robin / Neumann boundaries): one shared row, one tile column past the mesh
  if (use_tm && p.ov == 1 && dc == 1 && sw.stencil_variant == 2)  // A/B: x updated between the two reductions
    return launch_plan<true, 3, 6, 1, 1, false>(s, a, p, num_sms, sw);
  if (use_tm && p.ov == 1 && dc == 1 && sw.stencil_variant != 1) return launch_plan<true, 3, 6, 1, 1>(s, a, p, num_sms, sw);
  return use_tm ? launch_plan<true, 3, 6, -1, -1>(s, a, p, num_sms, sw) : launch_plan<false, 3, 6, -1, -1>(s, a, p, num_sms, sw);
}

}  // namespace msgpu
