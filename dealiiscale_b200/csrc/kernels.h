// Launchers of the sm_100a kernels of libmsgpu.  Host code in
// msgpu.cu calls these; each launcher returns the number of kernels it launched.
#pragma once
#include <cuda_runtime.h>

#include "vm_ops.h"

namespace msgpu {

// Geometry shared by all kernels.  Device ordering: node r = ix*m + iy, cell c = cx*nc + cy.
struct MeshDims {
  int nc;      // cells per axis
  int m;       // nodes per axis
  int n;       // m*m
  int ld;      // leading dimension of the stored diagonals: n rounded up to even (16-byte rows for bulk copies)
  int ncells;  // nc*nc
  int q;       // Gauss points per direction
  int P;       // 1-D evaluation points per system: nc*q quadrature abscissae, then m vertices
  double h;    // cell side
};

// Per-context tables produced by the slot / line kernels.
struct Tables {
  const double* slots;  // [N][n_slots]
  const double* T0;     // [N][n0][P]   y0-dependent nodes, point index p0 = cx*q + qx | nc*q + vertex
  const double* T1;     // [N][n1][P]   y1-dependent nodes, point index p1 = qy*nc + cy | nc*q + vertex
  int n_slots, n0, n1;
};

constexpr int CELL_STRIDE = 18;  // 10 local matrix entries, 4 load entries, 4 det-weighted basis integrals
constexpr int FACE_STRIDE = 7;   // Robin mass (aa, ab, bb), load (a, b), stretch weights (a, b)
constexpr int QTAB_STRIDE = 14; // see fe_tables.h build_qtab
constexpr int Q1D_ROWS = 6;     // see fe_tables.h build_q1d
constexpr int NDIAG = 5;         // stored diagonals: offsets 0, +1, +m-1, +m, +m+1

enum ProblemFlags : int {
  PF_ELLIPTIC = 0,
  PF_BIO = 1,
  PF_PARABOLIC = 2
};

// Interpreter kernels keep one VM register file per thread in dynamic shared memory (n_regs doubles x threads).
// vm_prepare_launch opts the kernel in above the 48 KB default and REFUSES programs that need more than an SM has
// (more than ~220 live temporaries: the launch would fail with "invalid argument"); a refused launcher launches
// nothing, returns 0, and vm_launch_error() tells why until vm_clear_launch_error().
bool vm_prepare_launch(const void* kernel, size_t smem_bytes, const char* what);
const char* vm_launch_error();
void vm_clear_launch_error();

// K0 tables ---------------------------------------------------------------------------------
int launch_slots(cudaStream_t s, const VmProgram& prog, const double* xy, int N, double t, double* slots, int n_slots);
int launch_lines(cudaStream_t s, const VmProgram& prog0, const VmProgram& prog1, const double* coord0,
                 const double* coord1, int P, int N, const double* slots, int n_slots, double* T0, int n0, double* T1,
                 int n1);

// K1 cell moments: for systems [k0, k0+nk) writes cell[(kl*CELL_STRIDE + e)*ncells + c].
// full: F00 F01 F10 F11 G per quadrature point; F / G: the two halves as separate programs.
// When F does not depend on y (or there is no mapping) the stiffness part is closed-form and only
// G is evaluated per quadrature point.
struct MomentPrograms {
  const VmProgram* full;
  const VmProgram* F;
  const VmProgram* G;
  int has_map;
  int jac_depends_on_y;
};
int launch_cell_moments(cudaStream_t s, const MomentPrograms& mp, MeshDims md, Tables tb, const double* qtab,
                        const double* qsum, int k0, int nk, double stiffness_scale, double* cell, int* error_flag);
// one-time upload of the per-quadrature-point tables into constant memory (q = 2, 3, 12)
int upload_constant_tables(int q, const double* qtab, const double* qsum, const double* q1d);

// K1 faces: for systems [k0,k0+nk) and side in {0:x=-1, 1:x=+1, 2:y=-1, 3:y=+1} writes
// face[((kl*4 + side)*FACE_STRIDE + e)*nc + f].  One program per side: outputs F00..F11, BC.
int launch_face_moments(cudaStream_t s, const VmProgram& prog, int side, MeshDims md, Tables tb, const double* q1d_pt,
                        const double* q1d_wt, int k0, int nk, int has_map, double* face);

int launch_face_moments_all(cudaStream_t s, const VmProgram* const prog[4], MeshDims md, Tables tb,
                            const double* q1d_pt, const double* q1d_wt, int k0, int nk, int has_map, double* face);

// K1 gather: cell (and face) arrays -> stencil diagonals + load vectors of systems [k0,k0+nk).
struct GatherArgs {
  MeshDims md;
  int problem;
  int k0, nk;
  const double* cell;   // chunk-local
  int uniform_stiffness;  // 1: F does not depend on y, so the 10 matrix entries and the 4 det-weighted basis
                          // integrals are the same in every cell of a system: they are read from cell 0 only
  const double* face;   // chunk-local or null
  double robin_left, robin_right;  // kappa_2, kappa_4 (bio)
  double* diag;         // [N][NDIAG][ld]
  double* b0;           // [N][n]   load that does not depend on the macro solution
  double* jw;           // [N][n]   elliptic: int phi_i det(F)   (multiplies u_k; also the bulk-integral weights)
  double* edge;         // [N][4][m] bio: left stretch weights, right stretch weights, left bc load, right bc load
};
int launch_gather(cudaStream_t s, const GatherArgs& a);

// K2 Dirichlet data + elimination (elliptic) and right-hand-side finalisation (all problems).
int launch_boundary_values(cudaStream_t s, const VmProgram& prog, MeshDims md, Tables tb, int N, double* bv);
struct RhsArgs {
  MeshDims md;
  int problem;
  int N;
  const double* u;      // [N]
  const double* w;      // [N] or null
  double coef_u, coef_w;  // bio: kappa_1, kappa_3 ; elliptic: 1, 0 ; parabolic: kappa, p_F
  double dt;              // parabolic time step
  double* diag;
  const double* b0;
  const double* jw;
  const double* edge;
  const double* bv;     // [N][n] Dirichlet values (elliptic)
  double* b;            // [N][n]
  double* x;            // [N][n] start vector (elliptic: written; bio: untouched)
};
int launch_rhs_and_dirichlet(cudaStream_t s, const RhsArgs& a);

// K4 batched CG --------------------------------------------------------------------------------
// A/B and test switches of the CG launchers, read from the MSGPU_* environment ONCE per context (msgpu_create,
// msgpu_reload_switches) instead of on every launch.
struct CgSwitches {
  int stencil = 1;         // MSGPU_CG_STENCIL=0: never take the class-compressed kernel
  int stencil_tmem = 1;    // MSGPU_STENCIL_TMEM=0: its all-register variant (one system per SM)
  int stencil_ctas = 0;    // MSGPU_STENCIL_CTAS=n: at most n CTAs of it per SM
  int stencil_variant = 0; // MSGPU_STENCIL_VARIANT=1: masks for any overlap / dead columns even where the counts are known; 2: x updated between the reductions (A/B)
  int force_global = 0;    // MSGPU_CG_VARIANT=global: the streaming kernel
  int wide = 0;            // MSGPU_CG_WIDE=1: 1024 threads x 5 rows instead of 512 x 9 (strided kernel)
  int strided = 0;         // MSGPU_CG_STRIDED=1: strided row ownership
  int tmem = -1;           // MSGPU_CG_TMEM: 0 = plain strip kernel, 5 = force the 512 x 9 shape
  int tmem_small = 1;      // MSGPU_CG_TMEM_SMALL=0: plain strip kernel for 1153..2304 rows
  int cluster = 1;         // MSGPU_CG_CLUSTER=0: stream the matrix instead of the cluster kernel
  int debug_launch = 0;    // MSGPU_DEBUG_LAUNCH: print launch shapes
  double ssor_omega = 1.0;  // MSGPU_SSOR_OMEGA in (0, 2)
};
CgSwitches cg_switches_from_env();

struct CgArgs {
  MeshDims md;
  int N;
  const double* diag;       // [N or 1][NDIAG][ld]
  long long diag_stride;    // NDIAG*ld, or 0 when every system has the same matrix
  const double* b;          // [N][n]
  double* x;                // [N][n] in: start vector, out: solution
  double tol;
  double tol2;              // set by launch_cg: largest t with sqrt(t) <= tol
  int max_it;
  int precond;    // 0 identity, 1 Jacobi, 2 SSOR (four-colour ordering; systems that fit one SM)
  double omega;   // SSOR relaxation parameter (deal.II default 1)
  int* iters;               // [N]
  double* res;              // [N]
  int* work_counter;        // device int, zeroed by the launcher
  int* fail_flag;           // device int: set to 1 if a system did not converge
  // Device-side dispatch between two CG kernels launched back to back: with `gate` set, a kernel returns at
  // once unless (*gate == 0) == (gate_run_if_zero != 0).  No host round trip decides which one does the work.
  const int* gate = nullptr;
  int gate_run_if_zero = 0;
  // Class-compressed matrices (kernels_cg_stencil.cu).  A system whose coefficients do not vary over the mesh
  // (F constant per system: every bio set, the linear elliptic maps, the parabolic driver) has only nine
  // distinct matrix rows - one per position class {first, inner, last}^2 - so the CG kernel keeps the stencil
  // in registers instead of the matrix in tensor memory, and two systems share an SM.
  const double* x_in = nullptr;  // class-compressed kernel only: start vectors read from here instead of x (null: x)
  double* stencil_ws = nullptr;  // [N or 1][STENCIL_TABLE] class tables, written by the compression kernel; null: off
  int* stencil_bad = nullptr;    // device int, raised when some system is NOT class-uniform
  int stencil_hint = 0;          // 0 unknown: compress + verify, both kernels launched behind the gate;
                                 // 1 verified before (matrix unchanged since): compress only; 2 known not uniform
  int dirichlet = 0;             // 1: boundary rows are eliminated Dirichlet rows (identity rows, zero couplings):
                                 //    the CG runs on the inner (m-2)^2 block, the boundary values stay as they are
  const CgSwitches* sw = nullptr;       // the context's switches; null: read the environment now
  // scratch of the streaming kernel k_cg_global (systems that fit neither an SM nor a cluster): owned by the caller's
  // context, grown on demand on the context's device, freed with it
  double** workspace = nullptr;
  size_t* workspace_bytes = nullptr;
};
constexpr int STENCIL_TABLE = 81;  // 9 position classes x 9 couplings
// returns launches; *variant = 0 shared-memory resident, 1 global-memory, 2 strip, 3 tensor memory, 4 cluster,
// 5 class-compressed stencil (decided on the device when stencil_hint == 0)
int launch_cg(cudaStream_t s, const CgArgs& a, int num_sms, int* variant);

// Class-compressed CG (kernels_cg_stencil.cu).  stencil_supported: is there a strip shape for this mesh;
// launch_stencil_compress: class tables (+ verification of every stored entry against them);
// launch_cg_stencil: the solve.  Both return the number of kernels launched, < 0 on error.
bool stencil_supported(const MeshDims& md, int dirichlet);
int launch_stencil_compress(cudaStream_t s, const CgArgs& a, int verify);
int launch_cg_stencil(cudaStream_t s, const CgArgs& a, int num_sms, const CgSwitches& sw);

// K5/K6 coupling integrals --------------------------------------------------------------------
struct CouplingArgs {
  MeshDims md;
  int problem;
  int N;
  const double* x;
  const double* jw;     // elliptic: [N][n]; parabolic: shared [n] lumped weights (stride 0)
  long long jw_stride;
  const double* edge;   // bio
  double kappa_left, kappa_right;  // bio: kappa_2, kappa_4
  double* c0;           // [N] (already offset to this rank's block)
  double* c1;
  // Peer-memory exchange (multi-GPU, comm.cu): when n_peers > 0 every integral is ALSO stored into the
  // gather buffers of all ranks (peer-mapped pointers, already offset to this rank's block) and the last
  // CTA to finish raises this rank's flag on every peer — compute and allgather in one kernel.
  int n_peers;
  double* peer_c0[8];
  double* peer_c1[8];
  unsigned int* peer_flag[8];
  unsigned int flag_value;
  int* done_counter;    // device int, zero between launches
};
int launch_coupling(cudaStream_t s, const CouplingArgs& a);

// K8: per-grid error / residual integrals and the point-evaluated load (parabolic initial data) ----
// cellred: [N][2][ncells] scratch; out2: [N][2] = per-grid sums of the two per-cell quantities
int launch_cell_errors(cudaStream_t s, const VmProgram& prog, MeshDims md, const double* slots, int n_slots, int qe,
                       const double* pt, const double* wt, const double* x, int N, double fd, double* cellred,
                       double* out2);
int launch_cell_residuals(cudaStream_t s, MeshDims md, const double* x, const double* xold, int N, double* cellred,
                          double* out2);
int launch_cell_point_load(cudaStream_t s, const VmProgram& prog, MeshDims md, const double* slots, int n_slots, int qe,
                           const double* pt, const double* wt, int k0, int nk, double* cell);

// y = M x for the shared symmetric stencil matrix (parabolic K7 building block) ----------------
int launch_shared_spmv(cudaStream_t s, MeshDims md, int N, const double* diag, const double* x, double* y);

}  // namespace msgpu
