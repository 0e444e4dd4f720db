// libmsgpu: context and C ABI (include/msgpu.h).  Host-side orchestration only — all arithmetic
// of the micro path runs in the kernels of kernels_assemble.cu / kernels_cg.cu / kernels_extra.cu.
//
// Mirrors the call sequence of the reference's MicroSolver:
//   setup()            src/elliptic-elliptic/micro.cpp:32-37   -> msgpu_setup
//   assemble_system()  :140-171                                -> msgpu_assemble
//   solve()            :175-182                                -> msgpu_solve
// and of MacroSolver::compute_microscopic_contribution (macro.cpp:162-166) -> msgpu_coupling.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/msgpu.h"
#include "ctx.h"
#include "expr_compile.h"
#include "fe_tables.h"
#include "kernels.h"

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

// a launcher refused an interpreter kernel (an expression with more live temporaries than an SM can hold)
#define CKVM()                                      \
  do {                                              \
    if (const char* e_ = vm_launch_error()) {       \
      ctx->err = e_;                                \
      vm_clear_launch_error();                      \
      return MSGPU_ERR_UNSUPPORTED;                 \
    }                                               \
  } while (0)

template <typename T>
int dev_alloc(msgpu_ctx* ctx, T** p, size_t count) {
  if (*p) {
    cudaFree(*p);
    *p = nullptr;
  }
  if (count == 0) return MSGPU_OK;
  CK(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  ctx->bytes_allocated += count * sizeof(T);
  return MSGPU_OK;
}

int fail(msgpu_ctx* ctx, int code, const std::string& msg) {
  ctx->err = msg;
  return code;
}

void free_all(msgpu_ctx* ctx) {
  void* ptrs[] = {ctx->d_xy,   ctx->d_slots, ctx->d_T0,  ctx->d_T1,   ctx->d_coord0, ctx->d_coord1, ctx->d_qtab,
                  ctx->d_q1pt, ctx->d_q1wt,  ctx->d_cell, ctx->d_face, ctx->d_diag,   ctx->d_b0,     ctx->d_jw,
                  ctx->d_edge, ctx->d_bv,    ctx->d_b,   ctx->d_x,    ctx->d_xold,   ctx->d_u,      ctx->d_w,
                  ctx->d_c,    ctx->d_res,   ctx->d_iters, ctx->d_flags, ctx->d_err,  ctx->d_mass,   ctx->d_macro,
                  ctx->d_lumped, ctx->d_cellred, ctx->d_eq_pt, ctx->d_eq_wt, ctx->d_stencil, ctx->d_cg_ws};
  for (void* p : ptrs)
    if (p) cudaFree(p);
}

struct PhaseTimer {
  msgpu_ctx* ctx;
  int phase;
  cudaStream_t s;
  cudaEvent_t a = nullptr, b = nullptr;
  PhaseTimer(msgpu_ctx* c, int ph, cudaStream_t stream = nullptr) : ctx(c), phase(ph), s(stream ? stream : c->stream) {
    if (!ctx->timing) return;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, s);
  }
  ~PhaseTimer() {
    if (!ctx->timing) return;
    cudaEventRecord(b, s);
    ctx->pending.push_back({phase, a, b});
  }
};

void drain_timers(msgpu_ctx* ctx) {
  for (auto& p : ctx->pending) {
    cudaEventSynchronize(p.b);
    float ms = 0;
    cudaEventElapsedTime(&ms, p.a, p.b);
    ctx->phase_ms[p.phase] += ms;
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  ctx->pending.clear();
}

std::map<std::string, double> const_map(const FnDef& f) {
  std::map<std::string, double> m;
  for (size_t i = 0; i < f.const_names.size(); ++i) m[f.const_names[i]] = f.const_values[i];
  return m;
}

// Build the expression DAG and all VM programs of this context.
int compile_programs(msgpu_ctx* ctx) {
  std::string err;
  const int rc = build_problem_programs(ctx->desc.problem, ctx->fn, ctx->pp, err);
  if (rc == 1) return fail(ctx, MSGPU_ERR_INVALID, err);
  if (rc == 2) return fail(ctx, MSGPU_ERR_PARSE, err);
  return MSGPU_OK;
}

Tables tables_of(const msgpu_ctx* ctx) {
  Tables tb;
  tb.slots = ctx->d_slots;
  tb.T0 = ctx->d_T0;
  tb.T1 = ctx->d_T1;
  tb.n_slots = ctx->pp.ps.n_slots();
  tb.n0 = ctx->pp.ps.n_line0();
  tb.n1 = ctx->pp.ps.n_line1();
  return tb;
}

// K0: (re)build slot and line tables for time t.
int build_tables(msgpu_ctx* ctx, double t) {
  PhaseTimer pt(ctx, PH_TABLES);
  const ProgramSet& ps = ctx->pp.ps;
  ctx->launches += launch_slots(ctx->stream, ps.slot_program(), ctx->d_xy, ctx->N, t, ctx->d_slots, ps.n_slots());
  ctx->launches += launch_lines(ctx->stream, ps.line0_program(), ps.line1_program(), ctx->d_coord0, ctx->d_coord1,
                                ctx->md.P, ctx->N, ctx->d_slots, ps.n_slots(), ctx->d_T0, ps.n_line0(), ctx->d_T1,
                                ps.n_line1());
  CKVM();
  CK(cudaGetLastError());
  ctx->tables_time = t;
  ctx->tables_valid = 1;
  return MSGPU_OK;
}

}  // namespace

namespace msgpu {
int build_tables_at(msgpu_ctx* ctx, double t) { return build_tables(ctx, t); }
int parabolic_setup(msgpu_ctx* ctx);                       // kernels_extra.cu / parabolic.cu
int parabolic_assemble(msgpu_ctx* ctx, double dt, double t);
int compute_errors(msgpu_ctx* ctx, double* l2, double* h1);
int compute_residuals(msgpu_ctx* ctx, double* l2);
int set_exact_solution(msgpu_ctx* ctx, double t);
int comm_allgather_coupling(msgpu_ctx* ctx, double* c0, double* c1);
int comm_block(const msgpu_ctx* ctx);
void comm_destroy(msgpu_ctx* ctx);
void comm_p2p_fill(msgpu_ctx* ctx, CouplingArgs& a);
int comm_p2p_collect(msgpu_ctx* ctx);
}  // namespace msgpu

extern "C" {

int msgpu_create(msgpu_ctx** out, const msgpu_desc* desc) {
  if (!out || !desc) return MSGPU_ERR_INVALID;
  *out = nullptr;
  if (desc->problem < 0 || desc->problem > 2 || desc->cells_per_axis < 1 || desc->q_degree < 1 ||
      desc->q_degree > 16 || desc->mesh_kind < 0 || desc->mesh_kind > 1)
    return MSGPU_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return MSGPU_ERR_CUDA;  // no CPU fallback
  msgpu_ctx* ctx = new msgpu_ctx();
  ctx->desc = *desc;
  if (desc->device >= 0) {
    if (cudaSetDevice(desc->device) != cudaSuccess) {
      delete ctx;
      return MSGPU_ERR_CUDA;
    }
    ctx->device = desc->device;
  } else {
    cudaGetDevice(&ctx->device);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) {
    delete ctx;
    return MSGPU_ERR_CUDA;
  }
  ctx->num_sms = prop.multiProcessorCount;
  // (above the default priority: msgpu_assemble_begin runs its kernels on a second stream of the lowest one)
  int least_prio = 0, greatest_prio = 0;
  cudaDeviceGetStreamPriorityRange(&least_prio, &greatest_prio);
  if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, greatest_prio) != cudaSuccess) {
    delete ctx;
    return MSGPU_ERR_CUDA;
  }
  ctx->own_stream = 1;
  ctx->cg_sw = cg_switches_from_env();
  {
    const char* e = std::getenv("MSGPU_REUSE_ASSEMBLY");
    ctx->reuse_assembly = e && e[0] == '1';
    const char* o = std::getenv("MSGPU_OVERLAP_ASSEMBLY");
    ctx->overlap_assembly = !(o && o[0] == '0');
  }
  try {
    ctx->ref_to_internal = reference_numbering(desc->mesh_kind, desc->cells_per_axis);
  } catch (const std::exception&) {
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MSGPU_ERR_INVALID;
  }
  *out = ctx;
  return MSGPU_OK;
}

// A pre-assembly that nobody has joined yet must not outlive the buffers or the inputs it works on.
static void retire_pre_assembly(msgpu_ctx* ctx) {
  if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
  ctx->pre_assembly_in_flight = 0;
}

void msgpu_destroy(msgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  retire_pre_assembly(ctx);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
  if (ctx->ev_pre) cudaEventDestroy(ctx->ev_pre);
  cudaStreamSynchronize(ctx->stream);
  drain_timers(ctx);
  comm_destroy(ctx);
  macro_release(ctx);
  norms_release(ctx);
  free_all(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* msgpu_last_error(const msgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int msgpu_set_stream(msgpu_ctx* ctx, void* cuda_stream) {
  if (!ctx) return MSGPU_ERR_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = 0;
  return MSGPU_OK;
}

int msgpu_reload_switches(msgpu_ctx* ctx) {
  if (!ctx) return MSGPU_ERR_INVALID;
  ctx->cg_sw = cg_switches_from_env();
  return MSGPU_OK;
}

int msgpu_synchronize(msgpu_ctx* ctx) {
  if (!ctx) return MSGPU_ERR_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  return MSGPU_OK;
}

int msgpu_set_grid_locations(msgpu_ctx* ctx, const double* xy, int n_grids) {
  if (ctx) retire_pre_assembly(ctx);
  if (!ctx || !xy || n_grids < 0) return MSGPU_ERR_INVALID;
  ctx->N = n_grids;
  ctx->xy.assign(xy, xy + 2 * static_cast<size_t>(n_grids));
  ctx->setup_done = 0;
  return MSGPU_OK;
}

int msgpu_set_function(msgpu_ctx* ctx, int slot, const char* expr, const char* const* const_names,
                       const double* const_values, int n_consts, int time_dependent) {
  if (ctx) retire_pre_assembly(ctx);
  if (!ctx || slot < 0 || slot >= MSGPU_FN_COUNT || !expr || n_consts < 0) return MSGPU_ERR_INVALID;
  FnDef& f = ctx->fn[slot];
  f.set = true;
  f.expr = expr;
  f.time_dep = time_dependent != 0;
  f.const_names.clear();
  f.const_values.clear();
  for (int i = 0; i < n_consts; ++i) {
    f.const_names.push_back(const_names[i]);
    f.const_values.push_back(const_values[i]);
  }
  ctx->setup_done = 0;
  // surface parse errors right away, as MultiscaleFunctionParser::initialize does
  try {
    ProgramSet probe;
    std::string text = f.expr;
    int ncomp = 1;
    {
      // count components (trailing ';' tolerated)
      int semi = 0;
      size_t last = std::string::npos;
      for (size_t i = 0; i < text.size(); ++i)
        if (text[i] == ';') {
          ++semi;
          last = i;
        }
      ncomp = semi + 1;
      if (last != std::string::npos && text.find_first_not_of(" \t\r\n", last + 1) == std::string::npos) ncomp = semi;
    }
    const int want = slot == MSGPU_FN_MAPPING ? 2 : slot == MSGPU_FN_JACOBIAN ? 4 : 1;
    if (ncomp != want)
      return fail(ctx, MSGPU_ERR_PARSE,
                  "expression <" + text + "> has " + std::to_string(ncomp) + " component(s), " + std::to_string(want) +
                      " expected");
    probe.parse_function(text, const_map(f), f.time_dep, want);
  } catch (const std::exception& e) {
    f.set = false;
    return fail(ctx, MSGPU_ERR_PARSE, e.what());
  }
  return MSGPU_OK;
}

int msgpu_set_scalars(msgpu_ctx* ctx, const double* values, int n) {
  if (!ctx || !values || n < 0 || n > MSGPU_MAX_SCALARS) return MSGPU_ERR_INVALID;
  retire_pre_assembly(ctx);
  for (int i = 0; i < n; ++i) ctx->scalars[i] = values[i];
  ctx->n_scalars = n;
  ctx->stencil_hint = 0;  // the Robin terms of the matrices change with the scalars
  ctx->assembly_cached = 0;
  return MSGPU_OK;
}

int msgpu_setup(msgpu_ctx* ctx) {
  if (!ctx) return MSGPU_ERR_INVALID;
  retire_pre_assembly(ctx);
  if (ctx->N <= 0) return fail(ctx, MSGPU_ERR_INVALID, "set_grid_locations must be called before setup");
  CK(cudaSetDevice(ctx->device));
  const int prob = ctx->desc.problem;
  if (prob == MSGPU_BIO && ctx->n_scalars < 5) return fail(ctx, MSGPU_ERR_INVALID, "bio needs 5 scalars");
  if (prob == MSGPU_PARABOLIC && ctx->n_scalars < 4) return fail(ctx, MSGPU_ERR_INVALID, "parabolic needs 4 scalars");
  int rc = compile_programs(ctx);
  if (rc) return rc;

  MeshDims& md = ctx->md;
  md.nc = ctx->desc.cells_per_axis;
  md.m = md.nc + 1;
  md.n = md.m * md.m;
  md.ld = (md.n + 1) & ~1;
  md.ncells = md.nc * md.nc;
  md.q = ctx->desc.q_degree;
  md.P = md.nc * md.q + md.m;
  md.h = 2.0 / md.nc;
  const int N = ctx->N;
  const size_t n = md.n;

  std::vector<double> pt, wt;
  gauss_legendre_unit(md.q, pt, wt);
  ctx->q_pt = pt;
  ctx->q_wt = wt;
  std::vector<double> qtab = build_qtab(md.q, pt, wt);
  std::vector<double> qsum(QTAB_STRIDE, 0.0);
  for (int q = 0; q < md.q * md.q; ++q)
    for (int i = 0; i < QTAB_STRIDE; ++i) qsum[i] += qtab[static_cast<size_t>(q) * QTAB_STRIDE + i];
  const std::vector<double> q1d = build_q1d(md.q, pt, wt);
  if (upload_constant_tables(md.q, qtab.data(), qsum.data(), q1d.data()))
    return fail(ctx, MSGPU_ERR_CUDA, "cannot upload the quadrature tables to constant memory");
  // Specialise the cell-moment kernel for this problem's expressions (elliptic and bio assemble every cycle, the
  // parabolic driver every time step).  Any failure leaves the interpreter kernels in charge and is kept in jit_log.
  ctx->jit_kernels = 0;
  ctx->jit_errors_ready = 0;
  ctx->jit_log.clear();
  {
    const char* jit_env = std::getenv("MSGPU_JIT");
    if (!(jit_env && jit_env[0] == '0')) {  // parabolic: prog_vol = prog_vol_G, no mapping, QGauss(2)
      const ProgramSet& ps = ctx->pp.ps;
      MomentPrograms mp;
      mp.full = &ps.point_program(ctx->pp.prog_vol);
      mp.F = ctx->pp.prog_vol_F >= 0 ? &ps.point_program(ctx->pp.prog_vol_F) : mp.full;
      mp.G = &ps.point_program(ctx->pp.prog_vol_G);
      mp.has_map = ctx->pp.has_map;
      mp.jac_depends_on_y = ctx->pp.jac_depends_on_y;
      bool unrolled = false;
      const std::string src = cell_moments_source(mp, md.q, qtab.data(), qsum.data(), &unrolled);
      if (src.empty()) {
        ctx->jit_log = "the volume program uses an instruction the code generator does not lower";
      } else if (jit_compile(src, "k_cell_moments_jit", &ctx->jit_moments, &ctx->jit_log) == 0) {
        ctx->jit_moments.unrolled_rows = unrolled;
        ctx->jit_kernels = 1;
      }
    } else if (jit_env && jit_env[0] == '0') {
      ctx->jit_log = "disabled by MSGPU_JIT=0";
    }
  }
  std::vector<double> c0(md.P), c1(md.P);
  for (int p = 0; p < md.nc * md.q; ++p) {
    const int cx = p / md.q, qx = p % md.q;
    c0[p] = (-1.0 + cx * md.h) + md.h * pt[qx];
    const int qy = p / md.nc, cy = p % md.nc;
    c1[p] = (-1.0 + cy * md.h) + md.h * pt[qy];
  }
  for (int v = 0; v < md.m; ++v) c0[md.nc * md.q + v] = c1[md.nc * md.q + v] = -1.0 + v * md.h;

  ctx->bytes_allocated = 0;
  // buffers that are allocated at first use and sized by N: a set-up with more grids must not inherit them
  for (double** p : {&ctx->d_cellred, &ctx->d_mass, &ctx->d_lumped})
    if (*p) {
      cudaFree(*p);
      *p = nullptr;
    }
  const ProgramSet& ps = ctx->pp.ps;
  if ((rc = dev_alloc(ctx, &ctx->d_xy, 2 * static_cast<size_t>(N)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_slots, static_cast<size_t>(N) * std::max(1, ps.n_slots())))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_T0, static_cast<size_t>(N) * std::max(1, ps.n_line0()) * md.P))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_T1, static_cast<size_t>(N) * std::max(1, ps.n_line1()) * md.P))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_coord0, static_cast<size_t>(md.P)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_coord1, static_cast<size_t>(md.P)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_qtab, qtab.size() + QTAB_STRIDE))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_q1pt, static_cast<size_t>(md.q)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_q1wt, static_cast<size_t>(md.q)))) return rc;
  // chunk of systems whose cell / face moments are in flight at once (<= 6 GiB of scratch of the 180 GB: the 4225
  // systems of `6 6` are one chunk - two launches per kernel instead of four, no second short tail)
  const size_t per_system = static_cast<size_t>(CELL_STRIDE) * md.ncells * sizeof(double);
  size_t chunk = (static_cast<size_t>(6) << 30) / per_system;
  if (const char* e = std::getenv("MSGPU_CHUNK")) chunk = static_cast<size_t>(std::atol(e));
  chunk = std::max<size_t>(1, std::min<size_t>(chunk, N));
  ctx->chunk = static_cast<int>(chunk);
  if ((rc = dev_alloc(ctx, &ctx->d_cell, chunk * CELL_STRIDE * md.ncells))) return rc;
  if (prob != MSGPU_ELLIPTIC)
    if ((rc = dev_alloc(ctx, &ctx->d_face, chunk * 4 * FACE_STRIDE * md.nc))) return rc;
  const bool shared_matrix = prob == MSGPU_PARABOLIC;
  if ((rc = dev_alloc(ctx, &ctx->d_diag, (shared_matrix ? 1 : static_cast<size_t>(N)) * NDIAG * md.ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_b0, N * n))) return rc;
  if (prob == MSGPU_ELLIPTIC) {
    if ((rc = dev_alloc(ctx, &ctx->d_jw, N * n))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_bv, N * n))) return rc;
    CK(cudaMemsetAsync(ctx->d_bv, 0, N * n * sizeof(double), ctx->stream));
  } else {
    if ((rc = dev_alloc(ctx, &ctx->d_edge, static_cast<size_t>(N) * 4 * md.m))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_xold, N * n))) return rc;
    CK(cudaMemsetAsync(ctx->d_xold, 0, N * n * sizeof(double), ctx->stream));
  }
  if ((rc = dev_alloc(ctx, &ctx->d_b, N * n))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_x, N * n))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_u, static_cast<size_t>(N)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_w, static_cast<size_t>(N)))) return rc;
  if (!ctx->comm_ready) {
    if ((rc = dev_alloc(ctx, &ctx->d_c, 2 * static_cast<size_t>(N)))) return rc;
    ctx->gather_capacity = 2 * static_cast<size_t>(N);
  }
  if ((rc = dev_alloc(ctx, &ctx->d_res, static_cast<size_t>(N)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_iters, static_cast<size_t>(N)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_flags, static_cast<size_t>(4)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_err, 2 * static_cast<size_t>(N)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_stencil, (shared_matrix ? 1 : static_cast<size_t>(N)) * STENCIL_TABLE))) return rc;
  ctx->stencil_hint = 0;

  CK(cudaMemcpyAsync(ctx->d_xy, ctx->xy.data(), 2 * sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_coord0, c0.data(), sizeof(double) * md.P, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_coord1, c1.data(), sizeof(double) * md.P, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_qtab, qtab.data(), sizeof(double) * qtab.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_qtab + qtab.size(), qsum.data(), sizeof(double) * QTAB_STRIDE, cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_q1pt, pt.data(), sizeof(double) * md.q, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_q1wt, wt.data(), sizeof(double) * md.q, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_x, 0, N * n * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_u, 0, N * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_w, 0, N * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // host vectors above go out of scope

  ctx->tables_valid = 0;
  ctx->assembly_cached = 0;
  if ((rc = build_tables(ctx, 0.0))) return rc;
  ctx->setup_done = 1;
  ctx->assembled = 0;
  if (prob == MSGPU_PARABOLIC) {
    if ((rc = parabolic_setup(ctx))) return rc;
  } else {
    // The dense equivalent of compute_pullback_objects (micro.cpp:75-94): assemble once, with the
    // macro values still zero, so that everything the coupling integrals need (det-weighted basis
    // integrals, boundary stretch weights, boundary loads) exists before the first macro solve —
    // the reference's first MacroSolver::run() integrates over the micro grids before any
    // MicroSolver::run() has happened (manager.cpp:53-58).
    if ((rc = msgpu_assemble(ctx))) return rc;
    // ... but the solutions themselves start as zero vectors (setup_scatter, micro.cpp:56-72); the
    // elliptic assembly has just written the Dirichlet values into them
    CK(cudaMemsetAsync(ctx->d_x, 0, N * n * sizeof(double), ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return MSGPU_OK;
}

int msgpu_num_grids(const msgpu_ctx* ctx) { return ctx ? ctx->N : 0; }
int msgpu_num_dofs(const msgpu_ctx* ctx) { return ctx ? ctx->md.n : 0; }

int msgpu_set_macro_solution(msgpu_ctx* ctx, const double* u, const double* w) {
  if (!ctx || !ctx->setup_done || !u) return MSGPU_ERR_INVALID;
  CK(cudaMemcpyAsync(ctx->d_u, u, sizeof(double) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  if (w) CK(cudaMemcpyAsync(ctx->d_w, w, sizeof(double) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the caller may reuse its buffers
  return MSGPU_OK;
}

// The part of the assembly that does not depend on the macro solution - cell and face moments, the gather into
// matrices, loads and weights - for all systems, on stream `st`.
static int assemble_macro_independent(msgpu_ctx* ctx, cudaStream_t st) {
  const MeshDims& md = ctx->md;
  const ProgramSet& ps = ctx->pp.ps;
  const int prob = ctx->desc.problem;
  const Tables tb = tables_of(ctx);
  const double d2 = prob == MSGPU_BIO ? ctx->scalars[4] : 1.0;
  for (int k0 = 0; k0 < ctx->N; k0 += ctx->chunk) {
    const int nk = std::min(ctx->chunk, ctx->N - k0);
    {
      PhaseTimer pt(ctx, PH_MOMENTS, st);
      MomentPrograms mp;
      mp.full = &ps.point_program(ctx->pp.prog_vol);
      mp.F = ctx->pp.prog_vol_F >= 0 ? &ps.point_program(ctx->pp.prog_vol_F) : mp.full;
      mp.G = &ps.point_program(ctx->pp.prog_vol_G);
      mp.has_map = ctx->pp.has_map;
      mp.jac_depends_on_y = ctx->pp.jac_depends_on_y;
      int launched = -1;
      if (ctx->jit_kernels)
        launched = launch_cell_moments_jit(st, ctx->jit_moments, md, tb, k0, nk, d2, ctx->d_cell,
                                           ctx->d_flags + FLAG_JACOBIAN);
      if (launched < 0)
        launched = launch_cell_moments(st, mp, md, tb, ctx->d_qtab,
                                       ctx->d_qtab + static_cast<size_t>(md.q) * md.q * QTAB_STRIDE, k0, nk, d2,
                                       ctx->d_cell, ctx->d_flags + FLAG_JACOBIAN);
      ctx->launches += launched;
      if (prob == MSGPU_BIO) {
        const VmProgram* faces[4];
        for (int s = 0; s < 4; ++s) faces[s] = &ps.point_program(ctx->pp.prog_face[s]);
        ctx->launches += launch_face_moments_all(st, faces, md, tb, ctx->d_q1pt, ctx->d_q1wt, k0, nk,
                                                 ctx->pp.has_map, ctx->d_face);
      }
    }
    {
      PhaseTimer pt(ctx, PH_GATHER, st);
      GatherArgs g{};
      g.uniform_stiffness = (!ctx->pp.has_map || !ctx->pp.jac_depends_on_y) ? 1 : 0;
      g.md = md;
      g.problem = prob;
      g.k0 = k0;
      g.nk = nk;
      g.cell = ctx->d_cell;
      g.face = prob == MSGPU_BIO ? ctx->d_face : nullptr;
      g.robin_left = prob == MSGPU_BIO ? ctx->scalars[1] : 0.0;
      g.robin_right = prob == MSGPU_BIO ? ctx->scalars[3] : 0.0;
      g.diag = ctx->d_diag;
      g.b0 = ctx->d_b0;
      g.jw = ctx->d_jw;
      g.edge = ctx->d_edge;
      ctx->launches += launch_gather(st, g);
    }
  }
  return MSGPU_OK;
}

int msgpu_assemble_begin(msgpu_ctx* ctx) {
  if (!ctx || !ctx->setup_done) return MSGPU_ERR_INVALID;
  // elliptic and bio (the parabolic step is one pass of its own), and not when the assembly is cached anyway
  if (ctx->desc.problem == MSGPU_PARABOLIC || !ctx->overlap_assembly || ctx->pre_assembly_in_flight) return MSGPU_OK;
  if (ctx->reuse_assembly && ctx->assembly_cached) return MSGPU_OK;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->tables_valid) {
    const int rc = build_tables(ctx, 0.0);
    if (rc) return rc;
  }
  if (!ctx->aux_stream) {
    // lowest priority: the CTAs of the (latency-bound, two-SM) macro solve on the main stream get the slots the
    // assembly kernels free before those kernels' own next CTAs do
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, least));
    CK(cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_pre, cudaEventDisableTiming));
  }
  // behind everything the main stream has been given so far (the last solve, coupling, stop test read what is
  // rewritten now), beside whatever it is given next
  CK(cudaEventRecord(ctx->ev_main, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_main, 0));
  int rc = assemble_macro_independent(ctx, ctx->aux_stream);
  if (rc == MSGPU_OK) {
    if (const char* e = vm_launch_error()) {  // a launcher refused an interpreter kernel
      ctx->err = e;
      vm_clear_launch_error();
      rc = MSGPU_ERR_UNSUPPORTED;
    } else if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ctx->ev_pre, ctx->aux_stream) != cudaSuccess) {
      ctx->err = "msgpu_assemble_begin: launch on the auxiliary stream failed";
      rc = MSGPU_ERR_CUDA;
    }
  }
  if (rc != MSGPU_OK) {
    // whatever did get launched must not run beside the full assembly the caller falls back to
    cudaStreamSynchronize(ctx->aux_stream);
    return rc;
  }
  ctx->pre_assembly_in_flight = 1;
  return MSGPU_OK;
}

int msgpu_assemble(msgpu_ctx* ctx) {
  if (!ctx || !ctx->setup_done) return MSGPU_ERR_INVALID;
  if (ctx->desc.problem == MSGPU_PARABOLIC)
    return fail(ctx, MSGPU_ERR_INVALID, "parabolic problems are advanced with msgpu_time_step");
  CK(cudaSetDevice(ctx->device));
  const MeshDims& md = ctx->md;
  const ProgramSet& ps = ctx->pp.ps;
  const int prob = ctx->desc.problem;
  const Tables tb = tables_of(ctx);
  if (!ctx->tables_valid) {
    int rc = build_tables(ctx, 0.0);
    if (rc) return rc;
  }
  const double d2 = prob == MSGPU_BIO ? ctx->scalars[4] : 1.0;
  // (bio only: the elliptic right-hand side eliminates the Dirichlet columns from the UNMASKED matrix, which a
  // second pass no longer has)
  const bool reuse = ctx->reuse_assembly && ctx->assembly_cached && prob == MSGPU_BIO;
  if (ctx->pre_assembly_in_flight) {
    // msgpu_assemble_begin has put the macro-independent part on the auxiliary stream: join it here
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pre, 0));
    ctx->pre_assembly_in_flight = 0;
  } else if (!reuse) {
    const int rc = assemble_macro_independent(ctx, ctx->stream);
    if (rc) return rc;
  }
  {
    PhaseTimer pt(ctx, PH_GATHER);
    if (prob == MSGPU_ELLIPTIC && !reuse)
      ctx->launches += launch_boundary_values(ctx->stream, ps.point_program(ctx->pp.prog_bv), md, tb, ctx->N, ctx->d_bv);
    RhsArgs r;
    r.md = md;
    r.problem = prob;
    r.N = ctx->N;
    r.u = ctx->d_u;
    r.w = ctx->d_w;
    r.coef_u = prob == MSGPU_BIO ? ctx->scalars[0] : 1.0;
    r.coef_w = prob == MSGPU_BIO ? ctx->scalars[2] : 0.0;
    r.diag = ctx->d_diag;
    r.b0 = ctx->d_b0;
    r.jw = ctx->d_jw;
    r.edge = ctx->d_edge;
    r.bv = ctx->d_bv;
    r.b = ctx->d_b;
    r.x = ctx->d_x;
    ctx->launches += launch_rhs_and_dirichlet(ctx->stream, r);
  }
  CKVM();
  CK(cudaGetLastError());
  ctx->assembled = 1;
  ctx->assembly_cached = 1;
  return MSGPU_OK;
}

int msgpu_solve(msgpu_ctx* ctx, double abs_tol, int max_it, int precond, int* iters, double* final_res) {
  if (!ctx || !ctx->setup_done || !ctx->assembled) return MSGPU_ERR_INVALID;
  if (precond != MSGPU_PRECOND_IDENTITY && precond != MSGPU_PRECOND_JACOBI && precond != MSGPU_PRECOND_SSOR)
    return fail(ctx, MSGPU_ERR_UNSUPPORTED, "preconditioner not supported (identity, Jacobi and SSOR are)");
  CK(cudaSetDevice(ctx->device));
  const MeshDims& md = ctx->md;
  const bool shared_matrix = ctx->desc.problem == MSGPU_PARABOLIC;
  // old_solutions = solutions (bio micro.cpp:300).  When the class-compressed kernel is known to do the solve it reads
  // its start vectors from one buffer and writes the solutions to the other: the two buffers change roles instead of
  // 2 x 143 MB being copied at `6 6`.
  bool swapped_in = false;
  if (ctx->desc.problem == MSGPU_BIO) {
    if (ctx->stencil_hint == 1 && precond == MSGPU_PRECOND_IDENTITY && ctx->cg_sw.stencil &&
        stencil_supported(md, 0)) {
      std::swap(ctx->d_x, ctx->d_xold);
      swapped_in = true;
    } else {
      CK(cudaMemcpyAsync(ctx->d_xold, ctx->d_x, sizeof(double) * ctx->N * md.n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  {
    PhaseTimer pt(ctx, PH_CG);
    CgArgs a;
    a.md = md;
    a.N = ctx->N;
    a.diag = ctx->d_diag;
    a.diag_stride = shared_matrix ? 0 : static_cast<long long>(NDIAG) * md.ld;
    a.b = ctx->d_b;
    a.x = ctx->d_x;
    a.x_in = swapped_in ? ctx->d_xold : nullptr;
    a.tol = abs_tol;
    a.max_it = max_it;
    a.precond = precond;
    a.omega = ctx->cg_sw.ssor_omega;  // PreconditionSSOR's default relaxation 1; MSGPU_SSOR_OMEGA in (0, 2) overrides
    a.sw = &ctx->cg_sw;
    a.workspace = &ctx->d_cg_ws;
    a.workspace_bytes = &ctx->cg_ws_bytes;
    a.iters = ctx->d_iters;
    a.res = ctx->d_res;
    a.work_counter = ctx->d_flags + FLAG_COUNTER;
    a.fail_flag = ctx->d_flags + FLAG_CGFAIL;
    a.stencil_ws = ctx->d_stencil;
    a.stencil_bad = ctx->d_flags + FLAG_STENCIL_BAD;
    a.stencil_hint = ctx->stencil_hint;
    a.dirichlet = ctx->desc.problem == MSGPU_ELLIPTIC ? 1 : 0;
    int variant = 0;
    int l = launch_cg(ctx->stream, a, ctx->num_sms, &variant);
    if (l == -2)
      return fail(ctx, MSGPU_ERR_UNSUPPORTED, "SSOR preconditioning needs a system that fits one SM (at most 67 nodes per axis)");
    if (l < 0) return fail(ctx, MSGPU_ERR_CUDA, "cannot allocate the CG workspace");
    ctx->launches += l;
    ctx->cg_variant = variant;
  }
  CK(cudaGetLastError());
  int flags[4];
  CK(cudaMemcpyAsync(flags, ctx->d_flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->h_iters.resize(ctx->N);
  CK(cudaMemcpyAsync(ctx->h_iters.data(), ctx->d_iters, sizeof(int) * ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  if (final_res)
    CK(cudaMemcpyAsync(final_res, ctx->d_res, sizeof(double) * ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if ((ctx->cg_variant & 0xff) == 5 && (ctx->cg_variant >> 8) != 0) {
    // both CG kernels were launched behind the device flag: it tells which one did the work, and the next
    // solves of this (unchanged) matrix launch only that one
    const bool uniform = flags[FLAG_STENCIL_BAD] == 0;
    ctx->stencil_hint = uniform ? 1 : 2;
    ctx->cg_variant = uniform ? 5 : (ctx->cg_variant >> 8);
  }
  long long total = 0;
  for (int v : ctx->h_iters) total += v;
  ctx->cg_iterations_total = total;
  if (iters) std::memcpy(iters, ctx->h_iters.data(), sizeof(int) * ctx->N);
  if (flags[FLAG_JACOBIAN]) {
    int zero = 0;
    cudaMemcpy(ctx->d_flags + FLAG_JACOBIAN, &zero, sizeof(int), cudaMemcpyHostToDevice);
    return fail(ctx, MSGPU_ERR_JACOBIAN, "Determinant of jacobian of mapping is not positive!");
  }
  if (flags[FLAG_CGFAIL]) return fail(ctx, MSGPU_ERR_NO_CONVERGENCE, "SolverControl::NoConvergence in a micro system");
  return MSGPU_OK;
}

int msgpu_time_step(msgpu_ctx* ctx, double dt, double t, double abs_tol, int max_it, int* iters) {
  if (!ctx || !ctx->setup_done || ctx->desc.problem != MSGPU_PARABOLIC) return MSGPU_ERR_INVALID;
  int rc = parabolic_assemble(ctx, dt, t);
  if (rc) return rc;
  ctx->assembled = 1;
  rc = msgpu_solve(ctx, abs_tol, max_it, MSGPU_PRECOND_IDENTITY, iters, nullptr);
  if (rc) return rc;
  // old_solutions = solutions (rho_solver.cpp:236)
  CK(cudaMemcpyAsync(ctx->d_xold, ctx->d_x, sizeof(double) * ctx->N * ctx->md.n, cudaMemcpyDeviceToDevice, ctx->stream));
  return MSGPU_OK;
}

int msgpu_coupling(msgpu_ctx* ctx, double* c0, double* c1) {
  if (!ctx || !ctx->setup_done) return MSGPU_ERR_INVALID;
  const int prob = ctx->desc.problem;
  if (prob == MSGPU_BIO && c0 && !c1) return fail(ctx, MSGPU_ERR_INVALID, "bio coupling needs both output arrays");
  CK(cudaSetDevice(ctx->device));
  const int ntot = ctx->comm_ready ? ctx->n_total : ctx->N;
  const int koff = ctx->comm_ready ? ctx->k_offset : 0;
  // second field (bio outflow integrals) starts after the padded first field
  const size_t field = ctx->comm_ready ? static_cast<size_t>(comm_block(ctx)) * ctx->n_ranks : static_cast<size_t>(ctx->N);
  {
    PhaseTimer pt(ctx, PH_COUPLING);
    CouplingArgs a;
    a.md = ctx->md;
    a.problem = prob;
    a.N = ctx->N;
    a.x = ctx->d_x;
    a.jw = prob == MSGPU_PARABOLIC ? ctx->d_lumped : ctx->d_jw;
    a.jw_stride = prob == MSGPU_PARABOLIC ? 0 : ctx->md.n;
    a.edge = ctx->d_edge;
    a.kappa_left = prob == MSGPU_BIO ? ctx->scalars[1] : 0.0;
    a.kappa_right = prob == MSGPU_BIO ? ctx->scalars[3] : 0.0;
    a.c0 = ctx->d_c + koff;          // written straight into this rank's block of the gather buffer
    a.c1 = ctx->d_c + field + koff;
    ctx->c_view = ctx->d_c;
    comm_p2p_fill(ctx, a);  // peer-memory exchange attached: the kernel also stores into every rank's buffer
    ctx->launches += launch_coupling(ctx->stream, a);
    CK(cudaGetLastError());
    if (ctx->p2p_ready) {
      int rc = comm_p2p_collect(ctx);
      if (rc) return rc;
    } else if (ctx->comm_ready) {
      int rc = comm_allgather_coupling(ctx, nullptr, nullptr);
      if (rc) return rc;
    }
  }
  if (!c0) return MSGPU_OK;
  CK(cudaMemcpyAsync(c0, ctx->c_view, sizeof(double) * ntot, cudaMemcpyDeviceToHost, ctx->stream));
  if (c1 && prob == MSGPU_BIO)
    CK(cudaMemcpyAsync(c1, ctx->c_view + field, sizeof(double) * ntot, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MSGPU_OK;
}

int msgpu_errors(msgpu_ctx* ctx, double* l2, double* h1) {
  if (!ctx || !ctx->setup_done || !l2) return MSGPU_ERR_INVALID;
  return compute_errors(ctx, l2, h1);
}
int msgpu_residuals(msgpu_ctx* ctx, double* l2) {
  if (!ctx || !ctx->setup_done || !l2) return MSGPU_ERR_INVALID;
  return compute_residuals(ctx, l2);
}
int msgpu_set_exact_solution(msgpu_ctx* ctx, double t) {
  if (!ctx || !ctx->setup_done) return MSGPU_ERR_INVALID;
  // the projection uses the right-hand sides as scratch: msgpu_solve is refused until the next assembly
  if (ctx->desc.problem != MSGPU_PARABOLIC) ctx->assembled = 0;
  return set_exact_solution(ctx, t);
}

static int copy_vectors_out(msgpu_ctx* ctx, const double* dev, int k0, int k1, double* out) {
  if (!ctx || !ctx->setup_done || !out || k0 < 0 || k1 > ctx->N || k0 > k1) return MSGPU_ERR_INVALID;
  const size_t n = ctx->md.n;
  std::vector<double> tmp((k1 - k0) * n);
  CK(cudaMemcpyAsync(tmp.data(), dev + k0 * n, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < k1 - k0; ++k)
    for (size_t i = 0; i < n; ++i) out[k * n + i] = tmp[k * n + ctx->ref_to_internal[i]];
  return MSGPU_OK;
}

int msgpu_get_solutions(msgpu_ctx* ctx, int k0, int k1, double* out) {
  return copy_vectors_out(ctx, ctx ? ctx->d_x : nullptr, k0, k1, out);
}
int msgpu_get_rhs(msgpu_ctx* ctx, int k0, int k1, double* out) {
  return copy_vectors_out(ctx, ctx ? ctx->d_b : nullptr, k0, k1, out);
}

int msgpu_set_solutions(msgpu_ctx* ctx, int k0, int k1, const double* in) {
  if (!ctx || !ctx->setup_done || !in || k0 < 0 || k1 > ctx->N || k0 > k1) return MSGPU_ERR_INVALID;
  const size_t n = ctx->md.n;
  std::vector<double> tmp((k1 - k0) * n);
  for (int k = 0; k < k1 - k0; ++k)
    for (size_t i = 0; i < n; ++i) tmp[k * n + ctx->ref_to_internal[i]] = in[k * n + i];
  CK(cudaMemcpyAsync(ctx->d_x + k0 * n, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MSGPU_OK;
}

int msgpu_zero_solutions(msgpu_ctx* ctx) {
  if (!ctx || !ctx->setup_done) return MSGPU_ERR_INVALID;
  CK(cudaMemsetAsync(ctx->d_x, 0, sizeof(double) * ctx->N * ctx->md.n, ctx->stream));
  return MSGPU_OK;
}

int msgpu_get_dof_permutation(msgpu_ctx* ctx, int* ref_to_internal) {
  if (!ctx || !ref_to_internal) return MSGPU_ERR_INVALID;
  std::copy(ctx->ref_to_internal.begin(), ctx->ref_to_internal.end(), ref_to_internal);
  return MSGPU_OK;
}

int msgpu_get_dof_points(msgpu_ctx* ctx, double* xy) {
  if (!ctx || !xy) return MSGPU_ERR_INVALID;
  const int m = ctx->desc.cells_per_axis + 1;
  const double h = 2.0 / ctx->desc.cells_per_axis;
  for (int i = 0; i < m * m; ++i) {
    const int r = ctx->ref_to_internal[i];
    xy[2 * i] = -1.0 + (r / m) * h;
    xy[2 * i + 1] = -1.0 + (r % m) * h;
  }
  return MSGPU_OK;
}

int msgpu_get_matrix_dense(msgpu_ctx* ctx, int k, double* out) {
  if (!ctx || !ctx->setup_done || !ctx->assembled || !out || k < 0 || k >= ctx->N) return MSGPU_ERR_INVALID;
  const MeshDims& md = ctx->md;
  const size_t n = md.n;
  std::vector<double> dg(static_cast<size_t>(NDIAG) * md.ld);
  const size_t base = ctx->desc.problem == MSGPU_PARABOLIC ? 0 : static_cast<size_t>(k) * NDIAG * md.ld;
  CK(cudaMemcpyAsync(dg.data(), ctx->d_diag + base, sizeof(double) * dg.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<int> internal_to_ref(n);
  for (size_t i = 0; i < n; ++i) internal_to_ref[ctx->ref_to_internal[i]] = static_cast<int>(i);
  std::fill(out, out + n * n, 0.0);
  const int off[NDIAG] = {0, 1, md.m - 1, md.m, md.m + 1};
  for (size_t r = 0; r < n; ++r)
    for (int d = 0; d < NDIAG; ++d) {
      const size_t c = r + off[d];
      if (c >= n) continue;
      const double v = dg[static_cast<size_t>(d) * md.ld + r];
      if (d > 0) {  // skip the structural zeros of wrapped offsets
        const int ix = static_cast<int>(r) / md.m, iy = static_cast<int>(r) % md.m;
        const int jx = static_cast<int>(c) / md.m, jy = static_cast<int>(c) % md.m;
        if (std::abs(jx - ix) > 1 || std::abs(jy - iy) > 1) continue;
      }
      const int rr = internal_to_ref[r], cc = internal_to_ref[c];
      out[rr * n + cc] = v;
      out[cc * n + rr] = v;
    }
  return MSGPU_OK;
}

int msgpu_enable_timing(msgpu_ctx* ctx, int on) {
  if (!ctx) return MSGPU_ERR_INVALID;
  ctx->timing = on;
  return MSGPU_OK;
}

const char* msgpu_jit_log(const msgpu_ctx* ctx) { return ctx ? ctx->jit_log.c_str() : "null context"; }

int msgpu_get_stats(msgpu_ctx* ctx, msgpu_stats* out, int reset) {
  if (!ctx || !out) return MSGPU_ERR_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  drain_timers(ctx);
  out->ms_tables = ctx->phase_ms[PH_TABLES];
  out->ms_moments = ctx->phase_ms[PH_MOMENTS];
  out->ms_gather = ctx->phase_ms[PH_GATHER];
  out->ms_cg = ctx->phase_ms[PH_CG];
  out->ms_coupling = ctx->phase_ms[PH_COUPLING];
  out->launches = ctx->launches;
  out->cg_iterations_total = ctx->cg_iterations_total;
  out->cg_variant = ctx->cg_variant;
  out->jit_kernels = ctx->jit_kernels;
  if (reset) {
    for (double& v : ctx->phase_ms) v = 0;
    ctx->launches = 0;
  }
  return MSGPU_OK;
}

}  // extern "C"
