// The opaque context behind msgpu_ctx*: owns every device buffer of one rank's block of micro
// systems (the memory MicroSolver owns in the reference: solutions, righthandsides,
// system_matrices — src/elliptic-elliptic/micro.h:95-171).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/msgpu.h"
#include "expr_compile.h"
#include "problem_programs.h"
#include "kernels.h"
#include "jit.h"

namespace msgpu {

enum Phase { PH_TABLES = 0, PH_MOMENTS, PH_GATHER, PH_CG, PH_COUPLING, PH_COUNT };
enum FlagSlot { FLAG_JACOBIAN = 0, FLAG_COUNTER = 1, FLAG_CGFAIL = 2, FLAG_STENCIL_BAD = 3 };

// Device-resident macro system(s), see macro.cu.
struct MacroDevice {
  int ready = 0, fields = 0;
  MeshDims md{};
  std::vector<int> ref2int;
  double *d_diag = nullptr, *d_C = nullptr, *d_b0 = nullptr, *d_b = nullptr, *d_x = nullptr, *d_all = nullptr,
         *d_res = nullptr;
  int *d_ref2int = nullptr, *d_int2ref = nullptr, *d_iters = nullptr, *d_flags = nullptr;
  int source[2] = {0, 1}, has_system[2] = {0, 0}, rhs_given[2] = {0, 0};
  // PiSolver's load (pi_solver.cpp:76-122): + int phi_i rho_h pi_h, a product of two finite-element functions
  int product[2] = {0, 0}, nonlinear[2] = {0, 0};
  double theta[2] = {0, 0};
  double* d_old = nullptr;          // [fields][n] the field's previous solution (internal ordering)
  unsigned char* d_dirichlet = nullptr;  // [fields][n] rows whose load is a boundary value
};

// The stop test on the device (kernels_norms.cu): the macro mesh in the caller's cell order and scratch.
struct NormDevice {
  int ready = 0, n_cells = 0, n_values = 0, q = 0;
  double h = 0.0;
  int4* d_cells = nullptr;
  double *d_values = nullptr, *d_cell_sq = nullptr, *d_out = nullptr, *d_pt = nullptr, *d_wt = nullptr;
};

struct PendingTimer {
  int phase;
  cudaEvent_t a, b;
};

void macro_release(msgpu_ctx* ctx);
void norms_release(msgpu_ctx* ctx);

}  // namespace msgpu

struct msgpu_ctx {
  msgpu_desc desc{};
  std::string err;
  int device = 0;
  int num_sms = 1;
  cudaStream_t stream = nullptr;
  int own_stream = 0;

  // problem definition
  int N = 0;
  std::vector<double> xy;
  msgpu::FnDef fn[MSGPU_FN_COUNT];
  double scalars[MSGPU_MAX_SCALARS] = {0};
  int n_scalars = 0;

  // compiled expressions
  msgpu::ProblemPrograms pp;

  // run-time specialised kernels (jit.h); jit_log keeps why a specialisation was not possible
  msgpu::JitKernel jit_moments, jit_errors;
  int jit_errors_ready = 0;
  int jit_kernels = 0;
  std::string jit_log;

  // mesh
  msgpu::MeshDims md{};
  std::vector<int> ref_to_internal;
  std::vector<double> q_pt, q_wt;
  int chunk = 1;

  // device memory
  double *d_xy = nullptr, *d_slots = nullptr, *d_T0 = nullptr, *d_T1 = nullptr;
  double *d_coord0 = nullptr, *d_coord1 = nullptr, *d_qtab = nullptr, *d_q1pt = nullptr, *d_q1wt = nullptr;
  double *d_cell = nullptr, *d_face = nullptr;
  double *d_diag = nullptr, *d_b0 = nullptr, *d_jw = nullptr, *d_edge = nullptr, *d_bv = nullptr;
  double *d_b = nullptr, *d_x = nullptr, *d_xold = nullptr, *d_u = nullptr, *d_w = nullptr;
  double *d_c = nullptr, *d_res = nullptr, *d_err = nullptr;
  double *d_mass = nullptr, *d_macro = nullptr, *d_lumped = nullptr;  // parabolic: shared mass stencil, lumped weights
  double *d_cellred = nullptr, *d_eq_pt = nullptr, *d_eq_wt = nullptr;  // error / residual scratch and quadrature
  int eq_degree = 0;
  double parabolic_dt = -1.0;
  int *d_iters = nullptr, *d_flags = nullptr;
  double* d_cg_ws = nullptr;    // scratch of the streaming CG kernel (grown on demand, this context's device)
  size_t cg_ws_bytes = 0;
  msgpu::CgSwitches cg_sw;      // MSGPU_* switches of the CG launchers, read at msgpu_create / msgpu_reload_switches
  double* d_stencil = nullptr;  // [N or 1][STENCIL_TABLE] class tables of class-uniform systems (kernels_cg_stencil.cu)
  int stencil_hint = 0;         // 0 not known yet, 1 verified for the current matrix, 2 known not class-uniform
  size_t bytes_allocated = 0;

  // state
  int setup_done = 0, assembled = 0, tables_valid = 0;
  // MSGPU_REUSE_ASSEMBLY=1 (opt-in, read at msgpu_create): matrices, macro-independent loads and boundary data do not
  // depend on the macro solution, so after the first msgpu_assemble of a set-up only the right-hand sides are
  // refreshed.  Off by default: the reference re-assembles in every cycle and so does the default path.
  int reuse_assembly = 0, assembly_cached = 0;
  // msgpu_assemble_begin: the macro-independent part of the (bio) assembly on a second stream, beside the macro solve
  int overlap_assembly = 1;          // MSGPU_OVERLAP_ASSEMBLY=0: msgpu_assemble_begin does nothing
  int pre_assembly_in_flight = 0;    // launched on aux_stream, not yet joined by msgpu_assemble
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_main = nullptr, ev_pre = nullptr;
  double tables_time = 0.0;
  std::vector<int> h_iters;
  long long cg_iterations_total = 0;
  int cg_variant = 0;

  // measurement
  int timing = 0;
  std::vector<msgpu::PendingTimer> pending;
  double phase_ms[msgpu::PH_COUNT] = {0};
  long long launches = 0;

  // macro system(s) on the device
  msgpu::MacroDevice macro;
  msgpu::NormDevice norms;

  // multi-GPU
  int comm_ready = 0, rank = 0, n_ranks = 1, n_total = 0, k_offset = 0;
  void* nccl_comm = nullptr;
  size_t gather_capacity = 0;  // doubles available in d_c / d_macro
  std::vector<int> rank_counts, rank_offsets;
  // peer-memory exchange (comm.cu): exchange buffer = data[2 parities][2 fields][padded] + flags[n_ranks]
  int p2p_ready = 0;
  unsigned int p2p_iter = 0;
  void* p2p_local = nullptr;
  void* p2p_peer[8] = {nullptr};
  int* d_p2p_done = nullptr;
  void* wait_value_fn = nullptr;
  const double* c_view = nullptr;  // where the gathered coupling vectors of the last msgpu_coupling live
};
