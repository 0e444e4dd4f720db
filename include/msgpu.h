/* msgpu.h — C ABI of the B200 micro-problem engine (libmsgpu.so).
 *
 * Drop-in boundary for the batched micro path of 0mar/dealiiscale.  The reference has no
 * FFI layer: Manager owns MacroSolver<2> and MicroSolver<2> by value and they exchange raw
 * pointers (reference: src/elliptic-elliptic/manager.h:17-21, src/tools/pde_data.h:76-90).
 * The boundary is therefore the public surface of MicroSolver / RhoSolver plus the
 * micro-reading part of MacroSolver / PiSolver; every entry point below names the
 * reference member it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; every host pointer is caller-owned and is copied
 *     in / out synchronously (the call returns after the copy has completed);
 *   - the context owns all device memory and one CUDA stream;
 *   - every function returns an msgpu_status; no C++ exception crosses the boundary;
 *     msgpu_last_error() gives the message of the last failure on that context;
 *   - a context is used by one host thread at a time (as in the reference);
 *   - micro systems ("grids") are numbered 0..N-1 in the order of the macro DoFs handed to
 *     msgpu_set_grid_locations(); micro DoFs are exchanged in the REFERENCE numbering
 *     (deal.II first-touch numbering of the chosen mesh kind), whatever the device layout is.
 *   - there is no CPU fallback: without a CUDA device msgpu_create() fails.
 */
#ifndef MSGPU_H
#define MSGPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msgpu_ctx msgpu_ctx;

typedef enum {
  MSGPU_OK = 0,
  MSGPU_ERR_INVALID = 1,        /* bad argument / call order                                   */
  MSGPU_ERR_PARSE = 2,          /* expression cannot be parsed (mu::ParserError in the reference) */
  MSGPU_ERR_JACOBIAN = 3,       /* det F <= 1e-4 somewhere (Assert in pde_data.cpp:182, mapping.cpp:20) */
  MSGPU_ERR_NO_CONVERGENCE = 4, /* SolverControl::NoConvergence                                   */
  MSGPU_ERR_CUDA = 5,
  MSGPU_ERR_COMM = 6,
  MSGPU_ERR_UNSUPPORTED = 7
} msgpu_status;

/* Which driver's micro problem. */
typedef enum {
  MSGPU_ELLIPTIC = 0,  /* src/elliptic-elliptic/micro.cpp : Dirichlet, matrix differs per grid      */
  MSGPU_BIO = 1,       /* src/bio-multiscale/micro.cpp    : Robin left/right, Neumann top/bottom    */
  MSGPU_PARABOLIC = 2  /* src/elliptic-parabolic/rho_solver.cpp : implicit Euler, one shared matrix */
} msgpu_problem;

typedef enum {
  MSGPU_MESH_REFINE_GLOBAL = 0, /* hyper_cube(-1,1)+refine_global(r): Z-order cells (micro.cpp:40-44)   */
  MSGPU_MESH_SUBDIVIDED = 1     /* subdivided_hyper_cube(n,-1,1): row-major cells (bio micro.cpp:74)     */
} msgpu_mesh_kind;

typedef struct {
  int problem;        /* msgpu_problem                                                   */
  int mesh_kind;      /* msgpu_mesh_kind                                                 */
  int cells_per_axis; /* 2^refine_level or n_cells                                       */
  int q_degree;       /* Gauss points per direction: 12 (elliptic, bio) or 2 (parabolic) */
  int device;         /* CUDA device ordinal, -1 = current device                        */
} msgpu_desc;

/* Function slots of msgpu_set_function (the parsed functions of src/tools/pde_data.h). */
typedef enum {
  MSGPU_FN_MAPPING = 0,   /* "mapping"      2 components, zeta(x,y)                                 */
  MSGPU_FN_JACOBIAN = 1,  /* "jac_mapping"  4 components, row-major F = grad_y zeta                 */
  MSGPU_FN_RHS = 2,       /* micro_rhs | bulk_rhs_v | micro_rhs(t)                                  */
  MSGPU_FN_BC = 3,        /* micro_bc (elliptic Dirichlet data, composed with the mapping)          */
  MSGPU_FN_SOLUTION = 4,  /* micro_solution | solution_v (composed with the mapping, except parabolic) */
  MSGPU_FN_BC_V1 = 5,     /* bio: inflow  (left,  y0=-1) | parabolic: micro_bc_robin               */
  MSGPU_FN_BC_V2 = 6,     /* bio: outflow (right, y0=+1) | parabolic: micro_bc_neumann             */
  MSGPU_FN_BC_V3 = 7,     /* bio: top     (y1=+1)                                                   */
  MSGPU_FN_BC_V4 = 8,     /* bio: bottom  (y1=-1)                                                   */
  MSGPU_FN_INIT = 9,      /* parabolic: init_rho                                                    */
  MSGPU_FN_COUNT = 10
} msgpu_function_slot;

/* Scalar slots of msgpu_set_scalars.
 *   bio       : kappa_1, kappa_2, kappa_3, kappa_4, D_2   (bio micro.cpp:186-190)
 *   parabolic : kappa, p_F, R, D                          (rho_solver.cpp:127-130)            */
#define MSGPU_MAX_SCALARS 8

/* The reference solves with PreconditionIdentity (micro.cpp:179); that is the parity default.  Jacobi and SSOR are
 * offered for ill-conditioned maps.  SSOR sweeps run in the four-colour ordering of the lattice (a 9-point stencil
 * never couples two nodes of one colour), relaxation 1; systems up to 67 nodes per axis. */
typedef enum { MSGPU_PRECOND_IDENTITY = 0, MSGPU_PRECOND_JACOBI = 1, MSGPU_PRECOND_SSOR = 2 } msgpu_precond;

/* ---- life cycle: MicroSolver(data, refine_level) / ~MicroSolver ------------------------- */
int msgpu_create(msgpu_ctx** ctx, const msgpu_desc* desc);
void msgpu_destroy(msgpu_ctx* ctx);
const char* msgpu_last_error(const msgpu_ctx* ctx);
/* Run on a caller-provided cudaStream_t instead of the context's own stream (NULL = default stream). */
int msgpu_set_stream(msgpu_ctx* ctx, void* cuda_stream);
int msgpu_synchronize(msgpu_ctx* ctx);
/* The MSGPU_* A/B switches of the CG launchers (MSGPU_CG_STENCIL, MSGPU_CG_VARIANT, MSGPU_CG_TMEM, ...) are read from
 * the environment once, in msgpu_create; this reads them again (tests and A/B scripts that flip a switch on a live
 * context).  No reference counterpart. */
int msgpu_reload_switches(msgpu_ctx* ctx);

/* ---- MicroSolver::set_grid_locations (micro.cpp:242-245) --------------------------------
 * xy[N][2]: the macro support points owned by this context (its shard of the macro DoFs). */
int msgpu_set_grid_locations(msgpu_ctx* ctx, const double* xy, int n_grids);

/* ---- MultiscaleData / BioData / TwoPressureData function initialisation ------------------
 * (pde_data.cpp:25-35, :84-107, :144-158).  expr: ';'-separated components.  constants: the
 * muParser constants the data class injects.  time_dependent: the variable list ends in ",t". */
int msgpu_set_function(msgpu_ctx* ctx, int slot, const char* expr, const char* const* const_names,
                       const double* const_values, int n_consts, int time_dependent);
int msgpu_set_scalars(msgpu_ctx* ctx, const double* values, int n);

/* ---- MicroSolver::setup (micro.cpp:32-37; bio :62-69; RhoSolver::setup rho_solver.cpp:28-32)
 * Compiles the expressions, builds mesh tables, allocates N systems, evaluates the x- and
 * line-dependent parts of the mapping (the dense equivalent of compute_pullback_objects). */
int msgpu_setup(msgpu_ctx* ctx);

int msgpu_num_grids(const msgpu_ctx* ctx);
int msgpu_num_dofs(const msgpu_ctx* ctx);

/* ---- MicroSolver::set_macro_solution (micro.cpp:97-100; bio :170-175; rho_solver.cpp:96-102)
 * u[N] (elliptic: macro solution; bio: sol_u; parabolic: pi), w[N] (bio: sol_w) or NULL.
 * The reference borrows a pointer and reads it at every run(); here the values are copied. */
int msgpu_set_macro_solution(msgpu_ctx* ctx, const double* u, const double* w);

/* ---- MicroSolver::assemble_system (micro.cpp:140-171) / bio local_assemble_system+copier
 * (bio micro.cpp:183-298).  Kernels: line tables -> cell moments -> stencil gather (+ faces,
 * + Dirichlet elimination).  Matrices, right-hand sides and start vectors stay on the device. */
int msgpu_assemble(msgpu_ctx* ctx);
/* Optional, elliptic and bio: starts the part of msgpu_assemble that does not depend on the macro solution (cell and face moments,
 * the gather into matrices / loads / weights) on a second stream and returns at once; the next msgpu_assemble joins it
 * and only adds the macro-dependent right-hand sides.  Lets a host run MacroSolver::assemble_system + solve
 * (macro.cpp:57-114, bio macro.cpp:100-202; msgpu_macro_solve) beside it - the reference does the two one after the
 * other (manager.cpp:53-58, bio manager.cpp:65-68), the results are the same bits.  Call it after the msgpu_coupling / stop-test calls of the
 * cycle before; a no-op for the parabolic problem, with MSGPU_OVERLAP_ASSEMBLY=0 and while MSGPU_REUSE_ASSEMBLY=1 has the
 * assembly cached. */
int msgpu_assemble_begin(msgpu_ctx* ctx);

/* ---- MicroSolver::solve (micro.cpp:175-182; bio :309-314): batched CG, deal.II recurrence,
 * absolute tolerance on the recursive residual.  iters / final_res: [N] or NULL.
 * Returns MSGPU_ERR_NO_CONVERGENCE if any system needs more than max_it steps. */
int msgpu_solve(msgpu_ctx* ctx, double abs_tol, int max_it, int precond, int* iters, double* final_res);

/* ---- RhoSolver::iterate(dt) at time t (rho_solver.cpp:232-237): assemble, solve, old = new. */
int msgpu_time_step(msgpu_ctx* ctx, double dt, double t, double abs_tol, int max_it, int* iters);

/* ---- MacroSolver::compute_microscopic_contribution
 * elliptic  c0[k] = int v_k J dy                      (macro.cpp:134-166)
 * bio       c0[k], c1[k] = inflow / outflow integrals (bio macro.cpp:254-306)
 * parabolic c0[k] = int rho_k dy                      (pi_solver.cpp:178-197)
 * With a communicator (msgpu_comm_init) the arrays have n_total entries and are allgathered.
 * c0 == NULL: compute (and allgather) on the device only, no copy to the host. */
int msgpu_coupling(msgpu_ctx* ctx, double* c0, double* c1);

/* ---- next tier: MicroSolver::compute_all_errors per-grid part (micro.cpp:189-205),
 * bio get_residual (bio micro.cpp:368-378).  l2[N], h1[N] per grid; the macro-level
 * aggregation (micro.cpp:206-214) is host work on N numbers. */
int msgpu_errors(msgpu_ctx* ctx, double* l2, double* h1);
int msgpu_residuals(msgpu_ctx* ctx, double* l2);

/* ---- MicroSolver::set_exact_solution (micro.cpp:224-234; bio micro.cpp:381-391; RhoSolver
 * rho_solver.cpp:353-359): solutions[k] = VectorTools::project(QGauss(q_degree | 3 parabolic),
 * exact solution at x_k) for every grid — a batched L2 projection (load vectors + CG on the shared
 * mass stencil).  t: time of the exact solution (parabolic only).  Overwrites the solutions and the
 * right-hand sides; the reference uses it for the micro-exact / micro-error output files
 * (manager.cpp:107-128, bio manager.cpp:156-181). */
int msgpu_set_exact_solution(msgpu_ctx* ctx, double t);

/* ---- public data of MicroSolver: solutions (micro.h:97), in reference DoF numbering -------- */
int msgpu_get_solutions(msgpu_ctx* ctx, int k0, int k1, double* out /* [k1-k0][n] */);
int msgpu_set_solutions(msgpu_ctx* ctx, int k0, int k1, const double* in);
/* solutions.at(k) = 0 for all k (micro.cpp:151), on the device. */
int msgpu_zero_solutions(msgpu_ctx* ctx);
int msgpu_get_rhs(msgpu_ctx* ctx, int k0, int k1, double* out /* [k1-k0][n] */);
/* ref_to_internal[n]: position of reference DoF i in the device ordering. */
int msgpu_get_dof_permutation(msgpu_ctx* ctx, int* ref_to_internal);
/* support point of every reference DoF, xy[n][2] (DoFTools::map_dofs_to_support_points) */
int msgpu_get_dof_points(msgpu_ctx* ctx, double* xy);
/* System matrix of grid k as a dense n x n array in reference numbering (inspection / tests). */
int msgpu_get_matrix_dense(msgpu_ctx* ctx, int k, double* out /* [n][n] */);

/* ---- multi-GPU: one context per rank, contiguous blocks of macro DoFs per rank -------------
 * (no reference call site: replaces the pointer hand-off of manager.cpp:25-26).
 * id: 128 bytes from msgpu_comm_unique_id on rank 0, distributed by the caller. */
int msgpu_comm_unique_id(char id[128]);
/* The partition every rank must use: rank r owns [k0,k1) = [r*B, min((r+1)*B, n_total)), B = ceil(n_total/n_ranks). */
int msgpu_partition(int n_total, int n_ranks, int rank, int* k0, int* k1);
int msgpu_comm_init(msgpu_ctx* ctx, const char id[128], int rank, int n_ranks, int n_total, int k_offset);
/* Peer-memory exchange over NVLink / NVSwitch (one process per GPU): after msgpu_comm_init (its id may
 * be NULL when NCCL is not wanted at all) every rank exports the handle of its exchange buffer, the host
 * distributes the n_ranks handles by any means, every rank attaches them.  From then on msgpu_coupling is
 * ONE kernel that computes the integrals and stores them into the gather buffers of all ranks, signalled
 * with a flag per peer and awaited with a stream memory operation; the NCCL allgather is gone. <= 8 ranks. */
int msgpu_comm_p2p_export(msgpu_ctx* ctx, unsigned char handle[64]);
int msgpu_comm_p2p_attach(msgpu_ctx* ctx, const unsigned char* handles /* [n_ranks][64] */);
/* Back to the NCCL allgather (also after a failed attach on this or another rank); a no-op when nothing is attached. */
int msgpu_comm_p2p_detach(msgpu_ctx* ctx);
/* Per-grid scalars of all ranks, e.g. the error / residual norms the stop test aggregates over the macro mesh
 * (micro.cpp:206-214): local[fields][msgpu_num_grids] -> all[fields][n_total] on every rank, fields <= 2.
 * One grouped ncclAllGather (needs the communicator); without msgpu_comm_init it is a plain copy. */
int msgpu_comm_allgather(msgpu_ctx* ctx, const double* local, double* all, int fields);
/* Broadcast of the macro solution from root; every rank keeps its own block (set_macro_solution). */
int msgpu_broadcast_macro(msgpu_ctx* ctx, double* u_all, double* w_all, int root);

/* ---- the macro system(s) on the device -------------------------------------
 * Replaces MacroSolver::assemble_system + solve (src/elliptic-elliptic/macro.cpp:57-114,
 * src/bio-multiscale/macro.cpp:100-202) and the solve of PiSolver (src/elliptic-parabolic/pi_solver.cpp:124-149)
 * for the cycles after the first: the macro matrix never changes and its load is affine in the micro
 * contribution, b(c) = b0 + C c.  The macro mesh is the mesh of the grid locations: it must have exactly
 * one DoF per micro system (all ranks' systems when a communicator is attached), in deal.II numbering.
 *   msgpu_macro_setup       1 or 2 fields on a (cells_per_axis)^2 mesh of the given kind
 *   msgpu_macro_set_system  per field: A (after MatrixTools::apply_boundary_values; symmetric), C and b0 as
 *                           CSR values on one 9-point pattern (rowstart/col in deal.II numbering, any
 *                           column order); `source` = which coupling vector c is (0: c0, 1: c1 of
 *                           msgpu_coupling); x0 = start vector (Dirichlet values in place), may be NULL
 *   msgpu_macro_set_rhs     a host-assembled right-hand side for the NEXT solve only (PiSolver: its
 *                           load is a product of two finite-element functions)
 *   msgpu_macro_set_product_load  PiSolver::assemble_system (pi_solver.cpp:76-122): from now on every solve of this
 *                           field adds  int phi_i rho_h pi_h  (QGauss(2); rho_h, pi_h = Q1 interpolants of the
 *                           micro masses - the first coupling vector - and of theta * the field's previous solution;
 *                           nonlinear: min(|m|, 1) and theta * min(|pi|, sqrt|pi|), pi_solver.cpp:62-73, :235-246) to
 *                           the right-hand side the host gave with msgpu_macro_set_rhs (the part that depends on t
 *                           only), on every row that is not flagged in dirichlet[n]; old0[n] = the previous solution
 *                           before the first solve (pi_solver.cpp:55: 5); afterwards it follows the solves
 *   msgpu_macro_solve       b = b0 + C c from the device-resident result of the last msgpu_coupling,
 *                           SolverCG as msgpu_solve (warm start), and, if adopt != 0, the solution
 *                           becomes the macro solution of this rank's micro systems (field 0 -> u,
 *                           field 1 -> w) without a host round trip; iters[n_fields] may be NULL      */
int msgpu_macro_setup(msgpu_ctx* ctx, int n_fields, int mesh_kind, int cells_per_axis);
int msgpu_macro_set_system(msgpu_ctx* ctx, int field, int source, const int* rowstart, const int* col,
                           const double* A, const double* C, const double* b0, const double* x0);
int msgpu_macro_set_rhs(msgpu_ctx* ctx, int field, const double* b);
int msgpu_macro_set_product_load(msgpu_ctx* ctx, int field, double theta, int nonlinear, const double* old0,
                                 const unsigned char* dirichlet);
int msgpu_macro_set_solution(msgpu_ctx* ctx, int field, const double* x);
int msgpu_macro_get_solution(msgpu_ctx* ctx, int field, double* x);
int msgpu_macro_solve(msgpu_ctx* ctx, double abs_tol, int max_it, int adopt, int* iters);

/* ---- the stop test on the device -----------------------------------------------------------
 * MicroSolver::compute_all_errors / compute_all_residuals (micro.cpp:186-215, bio micro.cpp:317-378) aggregate the
 * per-grid norms as a NODAL function on the macro mesh and take its L2 norm (integrate_difference against zero +
 * compute_global_error); MacroSolver::compute_residual (bio macro.cpp:230-245) does the same with a macro vector.
 *   msgpu_norm_setup         the macro mesh once: cell_dofs[n_cells][4] (vertex order (0,0),(1,0),(0,1),(1,1); indices
 *                            into vectors with one value per grid location, n_values of them), cell side h
 *   msgpu_norm_of_values     that norm of a host vector values[n_values], QGauss(q)
 *   msgpu_norm_of_residuals  ... of the per-grid ||v_k - v_k_old|| (as msgpu_residuals), computed, allgathered over
 *                            the ranks (NCCL) and aggregated without leaving the device
 *   msgpu_norm_of_errors     ... of the per-grid L2 and H1 errors (as msgpu_errors)
 * Cells, points and the sum over the cells keep the caller's order and the host's arithmetic (no FMA contraction):
 * the values equal those of the host loop they replace.  MSGPU_ERR_UNSUPPORTED: several ranks without an NCCL
 * communicator (ranks that share a device) - the caller aggregates on the host then.                          */
int msgpu_norm_setup(msgpu_ctx* ctx, const int* cell_dofs, int n_cells, int n_values, double h);
int msgpu_norm_of_values(msgpu_ctx* ctx, const double* values, int q, double* out);
int msgpu_norm_of_residuals(msgpu_ctx* ctx, int q, double* out);
int msgpu_norm_of_errors(msgpu_ctx* ctx, int q, double* l2, double* h1);

/* ---- measurement hooks ------------------------------------------------------------------- */
typedef struct {
  double ms_tables;    /* slot + line table kernels                    */
  double ms_moments;   /* cell / face moment kernels                   */
  double ms_gather;    /* stencil gather + boundary elimination        */
  double ms_cg;        /* batched CG                                   */
  double ms_coupling;  /* coupling integrals (+ collectives)           */
  long long launches;  /* kernels launched by this context so far      */
  long long cg_iterations_total; /* sum over systems of the last solve */
  int cg_variant;      /* 0 = shared-memory resident (strided rows), 1 = global-memory, 2 = resident (row strips),
                          3 = row strips with the matrix in tensor memory,
                          4 = one system per thread-block cluster (2-8 SMs; systems with more than 67 nodes per axis) */
  int jit_kernels;     /* assembly kernels specialised at run time (NVRTC) for this problem; 0 = the
                          bytecode interpreter runs at every quadrature point (MSGPU_JIT=0, or no NVRTC) */
} msgpu_stats;
int msgpu_enable_timing(msgpu_ctx* ctx, int on);
/* Why msgpu_stats.jit_kernels is 0 (empty string when the specialised kernels are in use). */
const char* msgpu_jit_log(const msgpu_ctx* ctx);
int msgpu_get_stats(msgpu_ctx* ctx, msgpu_stats* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MSGPU_H */
