// TEST-ONLY harness: drives the product's per-thread assembly bodies (assemble_bodies.cuh), the
// expression compiler and the VM sequentially on the CPU, so that the exact device code can be
// checked against the oracle in the CPU-only test tier.  It is NOT part of libmsgpu.so and is
// never used by the product path (which fails loudly without a GPU).  The CG used here is a plain
// host loop over the 5-diagonal format — it exists only to turn assembled systems into solutions
// for comparison; the product's CG is the CUDA kernel in kernels_cg.cu, tested under `-m gpu`.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../dealiiscale_b200/csrc/assemble_bodies.cuh"
#include "../../dealiiscale_b200/csrc/extra_bodies.cuh"
#include "../../dealiiscale_b200/csrc/fe_tables.h"
#include "../../dealiiscale_b200/csrc/jit.h"
#include "../../dealiiscale_b200/csrc/problem_programs.h"

using namespace msgpu;

namespace {
constexpr int TBE = 4;  // emulated "threads per block": register j of thread t lives at R[j*TBE + t]

struct Emul {
  int problem, mesh_kind, nc, q;
  FnDef fn[10];
  double scalars[8] = {0};
  ProblemPrograms pp;
  MeshDims md{};
  int N = 0;
  std::vector<double> xy, slots, T0, T1, coord0, coord1, qtab, q1pt, q1wt;
  std::vector<double> cell, face, diag, b0, jw, edge, bv, b, x, xold, mass, lumped;
  double par_dt = -1;
  std::vector<int> ref_to_internal;
  std::vector<int> iters;
  std::string err;
  int jac_flag = 0;
  int vector_vm = 1;  // 1: the vectorised moments body (what the product launches for q = 2, 3, 12)
  std::vector<double> qsum, q1d;
  std::vector<double> R;
  Tables tables() const {
    Tables tb;
    tb.slots = slots.data();
    tb.T0 = T0.data();
    tb.T1 = T1.data();
    tb.n_slots = pp.ps.n_slots();
    tb.n0 = pp.ps.n_line0();
    tb.n1 = pp.ps.n_line1();
    return tb;
  }
  double* regs(const VmProgram& p, int width = 1) {
    R.assign(static_cast<size_t>(std::max(8, p.n_regs)) * width * TBE, 0.0);
    return R.data();
  }
};
}  // namespace

extern "C" {

void* emul_create(int problem, int mesh_kind, int nc, int q) {
  Emul* e = new Emul();
  e->problem = problem;
  e->mesh_kind = mesh_kind;
  e->nc = nc;
  e->q = q;
  return e;
}
void emul_destroy(void* h) { delete static_cast<Emul*>(h); }
const char* emul_error(void* h) { return static_cast<Emul*>(h)->err.c_str(); }

int emul_set_function(void* h, int slot, const char* expr, const char* const* names, const double* vals, int n,
                      int tdep) {
  Emul* e = static_cast<Emul*>(h);
  FnDef& f = e->fn[slot];
  f.set = true;
  f.expr = expr;
  f.time_dep = tdep != 0;
  f.const_names.clear();
  f.const_values.clear();
  for (int i = 0; i < n; ++i) {
    f.const_names.push_back(names[i]);
    f.const_values.push_back(vals[i]);
  }
  return 0;
}
int emul_set_vector_mode(void* h, int on) {
  static_cast<Emul*>(h)->vector_vm = on;
  return 0;
}
int emul_moment_mode(void* h) {
  Emul* e = static_cast<Emul*>(h);
  return !e->pp.has_map ? MODE_NO_MAP : (e->pp.jac_depends_on_y ? MODE_GENERAL : MODE_CONST_F);
}
// Text of the run-time specialised cell-moment kernel for this problem (what the product hands to
// NVRTC).  Returns the length needed (including the terminating zero); 0 = not expressible.
int emul_jit_source(void* h, char* buf, int len) {
  Emul* e = static_cast<Emul*>(h);
  const ProgramSet& ps = e->pp.ps;
  MomentPrograms mp;
  mp.full = &ps.point_program(e->pp.prog_vol);
  mp.F = e->pp.prog_vol_F >= 0 ? &ps.point_program(e->pp.prog_vol_F) : mp.full;
  mp.G = &ps.point_program(e->pp.prog_vol_G);
  mp.has_map = e->pp.has_map;
  mp.jac_depends_on_y = e->pp.jac_depends_on_y;
  bool unrolled = false;
  const std::string src = cell_moments_source(mp, e->md.q, e->qtab.data(), e->qsum.data(), &unrolled);
  if (src.empty()) return 0;
  if (buf && len > 0) {
    std::strncpy(buf, src.c_str(), len - 1);
    buf[len - 1] = 0;
  }
  return static_cast<int>(src.size()) + 1;
}
// Raw table pointers for running that text on the host (tests only).
int emul_table_pointers(void* h, const double** slots, const double** T0, const double** T1, int* dims) {
  Emul* e = static_cast<Emul*>(h);
  const Tables tb = e->tables();
  *slots = tb.slots;
  *T0 = tb.T0;
  *T1 = tb.T1;
  dims[0] = tb.n_slots;
  dims[1] = tb.n0;
  dims[2] = tb.n1;
  dims[3] = e->md.nc;
  dims[4] = e->md.m;
  dims[5] = e->md.n;
  dims[6] = e->md.ld;
  dims[7] = e->md.ncells;
  dims[8] = e->md.q;
  dims[9] = e->md.P;
  return 0;
}
double emul_mesh_h(void* h) { return static_cast<Emul*>(h)->md.h; }
int emul_get_cell_scratch(void* h, double* out, long long count) {
  Emul* e = static_cast<Emul*>(h);
  if (count > static_cast<long long>(e->cell.size())) count = static_cast<long long>(e->cell.size());
  std::memcpy(out, e->cell.data(), sizeof(double) * count);
  return static_cast<int>(count);
}

int emul_set_scalars(void* h, const double* v, int n) {
  Emul* e = static_cast<Emul*>(h);
  for (int i = 0; i < n; ++i) e->scalars[i] = v[i];
  return 0;
}

int emul_setup(void* h, const double* xy, int N, double t) {
  Emul* e = static_cast<Emul*>(h);
  int rc = build_problem_programs(e->problem, e->fn, e->pp, e->err);
  if (rc) return rc;
  try {
    e->ref_to_internal = reference_numbering(e->mesh_kind, e->nc);
  } catch (const std::exception& ex) {
    e->err = ex.what();
    return 3;
  }
  MeshDims& md = e->md;
  md.nc = e->nc;
  md.m = md.nc + 1;
  md.n = md.m * md.m;
  md.ld = (md.n + 1) & ~1;
  md.ncells = md.nc * md.nc;
  md.q = e->q;
  md.P = md.nc * md.q + md.m;
  md.h = 2.0 / md.nc;
  e->N = N;
  e->xy.assign(xy, xy + 2 * N);
  gauss_legendre_unit(md.q, e->q1pt, e->q1wt);
  e->qtab = build_qtab(md.q, e->q1pt, e->q1wt);
  e->q1d = build_q1d(md.q, e->q1pt, e->q1wt);
  e->qsum.assign(QTAB_STRIDE, 0.0);
  for (int q = 0; q < md.q * md.q; ++q)
    for (int i = 0; i < QTAB_STRIDE; ++i) e->qsum[i] += e->qtab[static_cast<size_t>(q) * QTAB_STRIDE + i];
  e->coord0.resize(md.P);
  e->coord1.resize(md.P);
  for (int p = 0; p < md.nc * md.q; ++p) {
    const int cx = p / md.q, qx = p % md.q;
    e->coord0[p] = (-1.0 + cx * md.h) + md.h * e->q1pt[qx];
    const int qy = p / md.nc, cy = p % md.nc;
    e->coord1[p] = (-1.0 + cy * md.h) + md.h * e->q1pt[qy];
  }
  for (int v = 0; v < md.m; ++v) e->coord0[md.nc * md.q + v] = e->coord1[md.nc * md.q + v] = -1.0 + v * md.h;
  const ProgramSet& ps = e->pp.ps;
  e->slots.assign(static_cast<size_t>(N) * std::max(1, ps.n_slots()), 0.0);
  e->T0.assign(static_cast<size_t>(N) * std::max(1, ps.n_line0()) * md.P, 0.0);
  e->T1.assign(static_cast<size_t>(N) * std::max(1, ps.n_line1()) * md.P, 0.0);
  if (ps.n_slots() > 0)
    for (long long g = 0; g < N; ++g)
      body_slots(ps.slot_program(), e->xy.data(), N, t, e->slots.data(), ps.n_slots(), g, e->regs(ps.slot_program()), TBE);
  if (ps.n_line0() + ps.n_line1() > 0) {
    const VmProgram& big = ps.line0_program().n_regs > ps.line1_program().n_regs ? ps.line0_program() : ps.line1_program();
    for (long long g = 0; g < static_cast<long long>(N) * md.P; ++g)
      body_lines(ps.line0_program(), ps.line1_program(), e->coord0.data(), e->coord1.data(), md.P, N, e->slots.data(),
                 ps.n_slots(), e->T0.data(), ps.n_line0(), e->T1.data(), ps.n_line1(), g, e->regs(big), TBE);
  }
  const size_t n = md.n;
  e->cell.assign(static_cast<size_t>(N) * CELL_STRIDE * md.ncells, 0.0);
  e->face.assign(static_cast<size_t>(N) * 4 * FACE_STRIDE * md.nc, 0.0);
  e->diag.assign(static_cast<size_t>(N) * NDIAG * md.ld, 0.0);
  e->b0.assign(N * n, 0.0);
  e->jw.assign(N * n, 0.0);
  e->edge.assign(static_cast<size_t>(N) * 4 * md.m, 0.0);
  e->bv.assign(N * n, 0.0);
  e->b.assign(N * n, 0.0);
  e->x.assign(N * n, 0.0);
  e->xold.assign(N * n, 0.0);
  e->iters.assign(N, 0);
  return 0;
}

int emul_program_stats(void* h, int* n_slots, int* n0, int* n1, int* vol_len) {
  Emul* e = static_cast<Emul*>(h);
  *n_slots = e->pp.ps.n_slots();
  *n0 = e->pp.ps.n_line0();
  *n1 = e->pp.ps.n_line1();
  *vol_len = e->pp.ps.point_program(e->pp.prog_vol).n_code;
  return 0;
}

int emul_disassemble(void* h, int which, char* buf, int len) {
  Emul* e = static_cast<Emul*>(h);
  const ProgramSet& ps = e->pp.ps;
  std::string s = which == 0   ? ps.disassemble(ps.slot_program())
                  : which == 1 ? ps.disassemble(ps.line0_program())
                  : which == 2 ? ps.disassemble(ps.line1_program())
                               : ps.disassemble(ps.point_program(which - 3));
  std::strncpy(buf, s.c_str(), len - 1);
  buf[len - 1] = 0;
  return 0;
}

// The volume program's outputs at quadrature point (cx,cy,qx,qy) of system k, through the tables.
int emul_eval_volume_point(void* h, int k, int cx, int cy, int qx, int qy, double* out, int n_out) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  const VmProgram& prog = e->pp.ps.point_program(e->pp.prog_vol);
  Tables tb = e->tables();
  VmIo io;
  io.R = e->regs(prog);
  io.r_stride = TBE;
  io.U = tb.slots + static_cast<size_t>(k) * tb.n_slots;
  io.l0_stride = io.l1_stride = md.P;
  io.OUT = nullptr;
  io.out_stride = 0;
  io.L0 = tb.T0 + static_cast<size_t>(k) * tb.n0 * md.P + cx * md.q + qx;
  io.L1 = tb.T1 + static_cast<size_t>(k) * tb.n1 * md.P + qy * md.nc + cy;
  vm_run(prog, 0, io);
  for (int i = 0; i < n_out; ++i) out[i] = io.R[(prog.out0 + i) * TBE];
  return 0;
}

// The MODE_POINTS programs (exact solution / initial condition) at an arbitrary point.
int emul_eval_free_point(void* h, int which /*0 err, 1 init*/, int k, double y0, double y1, double* out) {
  Emul* e = static_cast<Emul*>(h);
  const int id = which == 0 ? e->pp.prog_err : e->pp.prog_init;
  if (id < 0) return 1;
  const VmProgram& prog = e->pp.ps.point_program(id);
  Tables tb = e->tables();
  VmIo io;
  io.R = e->regs(prog);
  io.r_stride = TBE;
  io.U = tb.slots + static_cast<size_t>(k) * tb.n_slots;
  io.L0 = io.L1 = nullptr;
  io.OUT = nullptr;
  io.l0_stride = io.l1_stride = io.out_stride = 0;
  io.R[0] = y0;
  io.R[TBE] = y1;
  vm_run(prog, 0, io);
  *out = io.R[prog.out0 * TBE];
  return 0;
}

int emul_assemble(void* h, const double* u, const double* w) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  const ProgramSet& ps = e->pp.ps;
  const int N = e->N, prob = e->problem;
  Tables tb = e->tables();
  const double d2 = prob == PF_BIO ? e->scalars[4] : 1.0;
  const VmProgram& vol = ps.point_program(e->pp.prog_vol);
  const VmProgram& volG = ps.point_program(e->pp.prog_vol_G);
  const VmProgram& volF = e->pp.prog_vol_F >= 0 ? ps.point_program(e->pp.prog_vol_F) : vol;
  const int mode = emul_moment_mode(e);
  const TabGlobal tab{e->qtab.data(), e->qsum.data(), e->q1d.data(), md.q};
  const long long work = static_cast<long long>(N) * md.ncells;
  auto run_vec = [&](auto Qc, auto Wc) {
    constexpr int Q = decltype(Qc)::value, W = decltype(Wc)::value;
    for (long long g = 0; g < work; ++g) {
      if (mode == MODE_GENERAL)
        body_cell_moments_vec<Q, W, MODE_GENERAL>(vol, vol, md, tb, tab, 0, N, d2, e->cell.data(), &e->jac_flag, g,
                                                  e->regs(vol, W), TBE);
      else if (mode == MODE_CONST_F)
        body_cell_moments_vec<Q, W, MODE_CONST_F>(volF, volG, md, tb, tab, 0, N, d2, e->cell.data(), &e->jac_flag, g,
                                                  e->regs(volF.n_regs > volG.n_regs * W ? volF : volG,
                                                          volF.n_regs > volG.n_regs * W ? 1 : W), TBE);
      else
        body_cell_moments_vec<Q, W, MODE_NO_MAP>(volG, volG, md, tb, tab, 0, N, d2, e->cell.data(), &e->jac_flag, g,
                                                 e->regs(volG, W), TBE);
    }
  };
  if (e->vector_vm && md.q == 12) {
    if (mode == MODE_GENERAL) run_vec(std::integral_constant<int, 12>(), std::integral_constant<int, 6>());
    else run_vec(std::integral_constant<int, 12>(), std::integral_constant<int, 12>());
  } else if (e->vector_vm && md.q == 3) {
    run_vec(std::integral_constant<int, 3>(), std::integral_constant<int, 3>());
  } else if (e->vector_vm && md.q == 2) {
    run_vec(std::integral_constant<int, 2>(), std::integral_constant<int, 2>());
  } else {
    const VmProgram& sp = e->pp.has_map ? vol : volG;
    for (long long g = 0; g < work; ++g)
      body_cell_moments(sp, md, tb, e->qtab.data(), 0, N, d2, e->pp.has_map, e->cell.data(), &e->jac_flag, g,
                        e->regs(sp), TBE);
  }
  if (prob != PF_ELLIPTIC)
    for (int s = 0; s < 4; ++s) {
      const VmProgram& fp = ps.point_program(e->pp.prog_face[s]);
      for (long long g = 0; g < static_cast<long long>(N) * md.nc; ++g)
        body_face_moments(fp, s, md, tb, e->q1pt.data(), e->q1wt.data(), 0, N, e->pp.has_map, e->face.data(), g,
                          e->regs(fp), TBE);
    }
  GatherArgs ga{};
  ga.uniform_stiffness = (!e->pp.has_map || !e->pp.jac_depends_on_y) ? 1 : 0;
  ga.md = md;
  ga.problem = prob;
  ga.k0 = 0;
  ga.nk = N;
  ga.cell = e->cell.data();
  ga.face = prob == PF_BIO ? e->face.data() : nullptr;
  ga.robin_left = prob == PF_BIO ? e->scalars[1] : 0.0;
  ga.robin_right = prob == PF_BIO ? e->scalars[3] : 0.0;
  ga.diag = e->diag.data();
  ga.b0 = e->b0.data();
  ga.jw = prob == PF_ELLIPTIC ? e->jw.data() : nullptr;
  ga.edge = e->edge.data();
  for (long long g = 0; g < static_cast<long long>(N) * md.n; ++g) body_gather(ga, g);
  if (prob == PF_ELLIPTIC) {
    const VmProgram& bp = ps.point_program(e->pp.prog_bv);
    for (long long g = 0; g < static_cast<long long>(N) * 4 * md.nc; ++g)
      body_boundary_values(bp, md, tb, N, e->bv.data(), g, e->regs(bp), TBE);
  }
  RhsArgs ra;
  ra.md = md;
  ra.problem = prob;
  ra.N = N;
  ra.u = u;
  ra.w = w;
  ra.coef_u = prob == PF_BIO ? e->scalars[0] : 1.0;
  ra.coef_w = prob == PF_BIO ? e->scalars[2] : 0.0;
  ra.diag = e->diag.data();
  ra.b0 = e->b0.data();
  ra.jw = e->jw.data();
  ra.edge = e->edge.data();
  ra.bv = e->bv.data();
  ra.b = e->b.data();
  ra.x = e->x.data();
  for (long long g = 0; g < static_cast<long long>(N) * md.n; ++g) body_rhs(ra, g);
  if (prob == PF_ELLIPTIC)
    for (long long g = 0; g < static_cast<long long>(N) * md.n; ++g) body_mask_dirichlet(md, N, e->diag.data(), g);
  return e->jac_flag ? 3 : 0;
}

static void spmv5(const MeshDims& md, const double* dg, const double* v, double* out) {
  const int off[4] = {1, md.m - 1, md.m, md.m + 1};
  for (int r = 0; r < md.n; ++r) {
    double s = dg[r] * v[r];
    for (int j = 0; j < 4; ++j) {
      const double* dj = dg + static_cast<size_t>(j + 1) * md.ld;
      if (r + off[j] < md.n) s += dj[r] * v[r + off[j]];
      if (r - off[j] >= 0) s += dj[r - off[j]] * v[r - off[j]];
    }
    out[r] = s;
  }
}

// test-only host CG (same recurrence as the kernel)
int emul_solve(void* h, double tol, int max_it, int* iters_out) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  const int n = md.n;
  int rc = 0;
  std::vector<double> g(n), d(n), hh(n);
  if (e->problem == PF_BIO) e->xold = e->x;
  for (int k = 0; k < e->N; ++k) {
    const double* dg = e->diag.data() + (e->problem == PF_PARABOLIC ? 0 : static_cast<size_t>(k) * NDIAG * md.ld);
    double* x = e->x.data() + static_cast<size_t>(k) * n;
    const double* b = e->b.data() + static_cast<size_t>(k) * n;
    spmv5(md, dg, x, hh.data());
    double rr = 0;
    for (int i = 0; i < n; ++i) {
      g[i] = hh[i] - b[i];
      rr += g[i] * g[i];
    }
    double res = std::sqrt(rr);
    int it = 0;
    bool ok = res <= tol;
    if (!ok) {
      for (int i = 0; i < n; ++i) d[i] = -g[i];
      double gh = res * res;
      for (;;) {
        ++it;
        spmv5(md, dg, d.data(), hh.data());
        double dh = 0;
        for (int i = 0; i < n; ++i) dh += d[i] * hh[i];
        const double alpha = gh / dh;
        rr = 0;
        for (int i = 0; i < n; ++i) {
          x[i] += alpha * d[i];
          g[i] += alpha * hh[i];
          rr += g[i] * g[i];
        }
        res = std::sqrt(rr);
        if (res <= tol) {
          ok = true;
          break;
        }
        if (it >= max_it || res != res) break;
        double beta = gh;
        gh = res * res;
        beta = gh / beta;
        for (int i = 0; i < n; ++i) d[i] = beta * d[i] - g[i];
      }
    }
    e->iters[k] = it;
    if (!ok) rc = 4;
  }
  if (iters_out) std::copy(e->iters.begin(), e->iters.end(), iters_out);
  return rc;
}

static void to_ref(const Emul* e, const double* internal, double* out) {
  for (int i = 0; i < e->md.n; ++i) out[i] = internal[e->ref_to_internal[i]];
}
int emul_get_rhs(void* h, int k, double* out) {
  Emul* e = static_cast<Emul*>(h);
  to_ref(e, e->b.data() + static_cast<size_t>(k) * e->md.n, out);
  return 0;
}
int emul_get_solution(void* h, int k, double* out) {
  Emul* e = static_cast<Emul*>(h);
  to_ref(e, e->x.data() + static_cast<size_t>(k) * e->md.n, out);
  return 0;
}
int emul_set_solution(void* h, int k, const double* in) {
  Emul* e = static_cast<Emul*>(h);
  for (int i = 0; i < e->md.n; ++i) e->x[static_cast<size_t>(k) * e->md.n + e->ref_to_internal[i]] = in[i];
  return 0;
}
int emul_get_matrix_dense(void* h, int k, double* out) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  const size_t n = md.n;
  const double* dg = e->diag.data() + (e->problem == PF_PARABOLIC ? 0 : static_cast<size_t>(k) * NDIAG * md.ld);
  std::vector<int> internal_to_ref(n);
  for (size_t i = 0; i < n; ++i) internal_to_ref[e->ref_to_internal[i]] = static_cast<int>(i);
  std::fill(out, out + n * n, 0.0);
  const int off[NDIAG] = {0, 1, md.m - 1, md.m, md.m + 1};
  for (size_t r = 0; r < n; ++r)
    for (int d = 0; d < NDIAG; ++d) {
      const size_t c = r + off[d];
      if (c >= n) continue;
      const int ix = static_cast<int>(r) / md.m, iy = static_cast<int>(r) % md.m;
      const int jx = static_cast<int>(c) / md.m, jy = static_cast<int>(c) % md.m;
      if (std::abs(jx - ix) > 1 || std::abs(jy - iy) > 1) continue;
      const double v = dg[static_cast<size_t>(d) * md.ld + r];
      out[internal_to_ref[r] * n + internal_to_ref[c]] = v;
      out[internal_to_ref[c] * n + internal_to_ref[r]] = v;
    }
  return 0;
}
// Wrapped offsets must hold exact zeros (the CG kernel relies on it instead of index tests).
int emul_check_wrapped_zeros(void* h) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  const int off[NDIAG] = {0, 1, md.m - 1, md.m, md.m + 1};
  int bad = 0;
  for (int k = 0; k < e->N; ++k)
    for (int r = 0; r < md.n; ++r)
      for (int d = 1; d < NDIAG; ++d) {
        const int c = r + off[d];
        const int ix = r / md.m, iy = r % md.m;
        const bool neighbour = c < md.n && std::abs(c / md.m - ix) <= 1 && std::abs(c % md.m - iy) <= 1;
        if (!neighbour && e->diag[(static_cast<size_t>(k) * NDIAG + d) * md.ld + r] != 0.0) ++bad;
      }
  return bad;
}
// same arithmetic as k_coupling (kernels_cg.cu), sequential
int emul_coupling(void* h, double* c0, double* c1) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  for (int k = 0; k < e->N; ++k) {
    const double* x = e->x.data() + static_cast<size_t>(k) * md.n;
    if (e->problem == PF_BIO) {
      const double* edge = e->edge.data() + static_cast<size_t>(k) * 4 * md.m;
      double su = 0, sw = 0;
      for (int iy = 0; iy < md.m; ++iy) {
        su += -e->scalars[1] * x[iy] * edge[iy] + edge[2 * md.m + iy];
        sw += -e->scalars[3] * x[(md.m - 1) * md.m + iy] * edge[md.m + iy] + edge[3 * md.m + iy];
      }
      c0[k] = su;
      c1[k] = sw;
    } else {
      const double* jw = e->jw.data() + static_cast<size_t>(k) * md.n;
      double s = 0;
      for (int r = 0; r < md.n; ++r) s += x[r] * jw[r];
      c0[k] = s;
    }
  }
  return 0;
}


// Text of the run-time specialised error kernel (exact-solution program), as emul_jit_source.
int emul_jit_errors_source(void* h, char* buf, int len) {
  Emul* e = static_cast<Emul*>(h);
  if (e->pp.prog_err < 0) return 0;
  const std::string src = cell_errors_source(e->pp.ps.point_program(e->pp.prog_err));
  if (src.empty()) return 0;
  if (buf && len > 0) {
    std::strncpy(buf, src.c_str(), len - 1);
    buf[len - 1] = 0;
  }
  return static_cast<int>(src.size()) + 1;
}
// Per-cell error integrals of the interpreter body: out[(k*2 + {0,1})*ncells + c].
int emul_cell_errors(void* h, double* out) {
  Emul* e = static_cast<Emul*>(h);
  if (e->pp.prog_err < 0) return 1;
  const MeshDims& md = e->md;
  const int qe = e->problem == PF_PARABOLIC ? 3 : md.q;
  std::vector<double> pt, wt;
  gauss_legendre_unit(qe, pt, wt);
  const VmProgram& prog = e->pp.ps.point_program(e->pp.prog_err);
  for (long long g = 0; g < static_cast<long long>(e->N) * md.ncells; ++g)
    body_cell_errors(prog, md, e->slots.data(), e->pp.ps.n_slots(), qe, pt.data(), wt.data(), e->x.data(), e->N, 1e-8,
                     out, g, e->regs(prog), TBE);
  return 0;
}
const double* emul_solution_pointer(void* h) { return static_cast<Emul*>(h)->x.data(); }

// ---- errors / residuals (same bodies as kernels_extra.cu, sequential) ---------------------------
int emul_errors(void* h, double* l2, double* h1) {
  Emul* e = static_cast<Emul*>(h);
  if (e->pp.prog_err < 0) return 1;
  const MeshDims& md = e->md;
  const int qe = e->problem == PF_PARABOLIC ? 3 : md.q;
  std::vector<double> pt, wt;
  gauss_legendre_unit(qe, pt, wt);
  const VmProgram& prog = e->pp.ps.point_program(e->pp.prog_err);
  std::vector<double> red(static_cast<size_t>(e->N) * 2 * md.ncells);
  for (long long g = 0; g < static_cast<long long>(e->N) * md.ncells; ++g)
    body_cell_errors(prog, md, e->slots.data(), e->pp.ps.n_slots(), qe, pt.data(), wt.data(), e->x.data(), e->N, 1e-8,
                     red.data(), g, e->regs(prog), TBE);
  for (int k = 0; k < e->N; ++k) {
    double a = 0, b = 0;
    for (int c = 0; c < md.ncells; ++c) {
      a += red[(static_cast<size_t>(k) * 2) * md.ncells + c];
      b += red[(static_cast<size_t>(k) * 2 + 1) * md.ncells + c];
    }
    l2[k] = std::sqrt(a);
    h1[k] = std::sqrt(e->problem == PF_PARABOLIC ? a + b : b);
  }
  return 0;
}
int emul_residuals(void* h, double* l2) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  std::vector<double> red(static_cast<size_t>(e->N) * 2 * md.ncells);
  for (long long g = 0; g < static_cast<long long>(e->N) * md.ncells; ++g)
    body_cell_residual(md, e->x.data(), e->xold.data(), e->N, red.data(), g);
  for (int k = 0; k < e->N; ++k) {
    double a = 0;
    for (int c = 0; c < md.ncells; ++c) a += red[(static_cast<size_t>(k) * 2) * md.ncells + c];
    l2[k] = std::sqrt(a);
  }
  return 0;
}

// ---- parabolic: mirrors parabolic_setup / parabolic_assemble of kernels_extra.cu ---------------
static void rebuild_tables(Emul* e, double t) {
  const MeshDims& md = e->md;
  const ProgramSet& ps = e->pp.ps;
  const int N = e->N;
  if (ps.n_slots() > 0)
    for (long long g = 0; g < N; ++g)
      body_slots(ps.slot_program(), e->xy.data(), N, t, e->slots.data(), ps.n_slots(), g, e->regs(ps.slot_program()), TBE);
  if (ps.n_line0() + ps.n_line1() > 0) {
    const VmProgram& big = ps.line0_program().n_regs > ps.line1_program().n_regs ? ps.line0_program() : ps.line1_program();
    for (long long g = 0; g < static_cast<long long>(N) * md.P; ++g)
      body_lines(ps.line0_program(), ps.line1_program(), e->coord0.data(), e->coord1.data(), md.P, N, e->slots.data(),
                 ps.n_slots(), e->T0.data(), ps.n_line0(), e->T1.data(), ps.n_line1(), g, e->regs(big), TBE);
  }
}
int emul_parabolic_init(void* h) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  const int N = e->N;
  std::vector<double> sys;
  build_stencils(md, 0.0, e->scalars[3], e->scalars[0], e->scalars[2], e->mass, sys, e->lumped);
  if (e->pp.prog_init < 0) return 0;
  std::vector<double> pt, wt;
  gauss_legendre_unit(3, pt, wt);
  const VmProgram& prog = e->pp.ps.point_program(e->pp.prog_init);
  for (long long g = 0; g < static_cast<long long>(N) * md.ncells; ++g)
    body_cell_point_load(prog, md, e->slots.data(), e->pp.ps.n_slots(), 3, pt.data(), wt.data(), 0, N, e->cell.data(), g,
                         e->regs(prog), TBE);
  GatherArgs ga{};
  ga.md = md;
  ga.problem = PF_PARABOLIC;
  ga.k0 = 0;
  ga.nk = N;
  ga.cell = e->cell.data();
  ga.face = nullptr;
  ga.diag = nullptr;
  ga.b0 = e->b.data();
  ga.jw = nullptr;
  ga.edge = e->edge.data();
  for (long long g = 0; g < static_cast<long long>(N) * md.n; ++g) body_gather(ga, g);
  e->diag = e->mass;
  std::fill(e->x.begin(), e->x.end(), 0.0);
  emul_solve(h, 1e-14, 10 * md.n + 100, nullptr);
  e->xold = e->x;
  return 0;
}
int emul_parabolic_step(void* h, const double* u, double dt, double t, double tol, int max_it, int* iters) {
  Emul* e = static_cast<Emul*>(h);
  const MeshDims& md = e->md;
  const ProgramSet& ps = e->pp.ps;
  const int N = e->N;
  const double kappa = e->scalars[0], p_F = e->scalars[1], R = e->scalars[2], D = e->scalars[3];
  std::vector<double> mass;
  build_stencils(md, dt, D, kappa, R, mass, e->diag, e->lumped);
  rebuild_tables(e, t);
  Tables tb = e->tables();
  const VmProgram& volG = ps.point_program(e->pp.prog_vol_G);
  const TabGlobal tab{e->qtab.data(), e->qsum.data(), e->q1d.data(), md.q};
  for (long long g = 0; g < static_cast<long long>(N) * md.ncells; ++g) {
    if (md.q == 2)
      body_cell_moments_vec<2, 2, MODE_NO_MAP>(volG, volG, md, tb, tab, 0, N, 1.0, e->cell.data(), &e->jac_flag, g,
                                               e->regs(volG, 2), TBE);
    else
      body_cell_moments(volG, md, tb, e->qtab.data(), 0, N, 1.0, 0, e->cell.data(), &e->jac_flag, g, e->regs(volG), TBE);
  }
  for (int s = 0; s < 4; ++s) {
    const VmProgram& fp = ps.point_program(e->pp.prog_face[s]);
    for (long long g = 0; g < static_cast<long long>(N) * md.nc; ++g)
      body_face_moments(fp, s, md, tb, e->q1pt.data(), e->q1wt.data(), 0, N, 0, e->face.data(), g, e->regs(fp), TBE);
  }
  GatherArgs ga{};
  ga.md = md;
  ga.problem = PF_PARABOLIC;
  ga.k0 = 0;
  ga.nk = N;
  ga.cell = e->cell.data();
  ga.face = e->face.data();
  ga.diag = nullptr;
  ga.b0 = e->b0.data();
  ga.jw = nullptr;
  ga.edge = e->edge.data();
  for (long long g = 0; g < static_cast<long long>(N) * md.n; ++g) body_gather(ga, g);
  for (int k = 0; k < N; ++k)
    spmv5(md, e->mass.data(), e->xold.data() + static_cast<size_t>(k) * md.n, e->b.data() + static_cast<size_t>(k) * md.n);
  RhsArgs ra{};
  ra.md = md;
  ra.problem = PF_PARABOLIC;
  ra.N = N;
  ra.u = u;
  ra.w = u;
  ra.coef_u = kappa;
  ra.coef_w = p_F;
  ra.dt = dt;
  ra.diag = e->diag.data();
  ra.b0 = e->b0.data();
  ra.edge = e->edge.data();
  ra.b = e->b.data();
  ra.x = e->x.data();
  for (long long g = 0; g < static_cast<long long>(N) * md.n; ++g) body_rhs(ra, g);
  int rc = emul_solve(h, tol, max_it, iters);
  e->xold = e->x;
  return rc;
}
int emul_mass_integrals(void* h, double* c0) {
  Emul* e = static_cast<Emul*>(h);
  for (int k = 0; k < e->N; ++k) {
    double s = 0;
    for (int r = 0; r < e->md.n; ++r) s += e->x[static_cast<size_t>(k) * e->md.n + r] * e->lumped[r];
    c0[k] = s;
  }
  return 0;
}

int emul_reference_numbering(int mesh_kind, int nc, int* out) {
  try {
    const std::vector<int> v = reference_numbering(mesh_kind, nc);
    std::copy(v.begin(), v.end(), out);
    return 0;
  } catch (const std::exception&) {
    return 1;
  }
}

}  // extern "C"
