// Bytecode -> CUDA C lowering and the text of the specialised cell-moment kernel (see jit.h).
// Pure host code without CUDA calls: tests/native compiles it too and feeds the result to NVRTC.
#include "jit.h"

#include "fe_tables.h"

#include <cstdio>
#include <cstring>
#include <sstream>
#include <vector>

namespace msgpu {

namespace {

// ------------------------------------------------------------------------------ lowering
std::string dbl(double v) {
  long long bits;
  std::memcpy(&bits, &v, sizeof bits);
  char buf[64];
  std::snprintf(buf, sizeof buf, "__longlong_as_double(0x%016llxLL)", static_cast<unsigned long long>(bits));
  return buf;
}

const char* unary_fn(uint32_t op) {
  switch (op) {
    case OP_SIN: return "sin";
    case OP_COS: return "cos";
    case OP_TAN: return "tan";
    case OP_ASIN: return "asin";
    case OP_ACOS: return "acos";
    case OP_ATAN: return "atan";
    case OP_SINH: return "sinh";
    case OP_COSH: return "cosh";
    case OP_TANH: return "tanh";
    case OP_ASINH: return "asinh";
    case OP_ACOSH: return "acosh";
    case OP_ATANH: return "atanh";
    case OP_LOG2: return "log2";
    case OP_LOG10: return "log10";
    case OP_LOG: return "log";
    case OP_EXP: return "exp";
    case OP_SQRT: return "sqrt";
    case OP_ABS: return "fabs";
    case OP_CEIL: return "ceil";
    case OP_FLOOR: return "floor";
    case OP_ERFC: return "erfc";
    default: return nullptr;
  }
}

}  // namespace

std::string vm_to_cuda(const VmProgram& prog, const std::string& name, int W) {
  std::ostringstream o;
  const int n_regs = prog.n_regs < 1 ? 1 : prog.n_regs;
  o << "__device__ __forceinline__ void " << name
    << "(const double* __restrict__ U, const double* __restrict__ L0, int l0s, const double* __restrict__ L1, int l1s, "
       "double (&R)["
    << n_regs << "][" << W << "]) {\n";
  o << "  double A[" << W << "];\n";
  const std::string lanes = "  _Pragma(\"unroll\") for (int e = 0; e < " + std::to_string(W) + "; ++e) ";
  o << lanes << "A[e] = 0.0;\n";
  for (int pc = 0; pc < prog.n_code; ++pc) {
    const uint32_t ins = prog.code[pc];
    const uint32_t op = ins & 0xffu;
    if (op == OP_END) break;
    const uint32_t kind = (ins >> 8) & 7u;
    const uint32_t idx = (ins >> 11) & 0x3ffu;
    if (op <= OP_LAST_BINARY) {
      std::string b;
      switch (kind) {
        case VK_CONST: b = dbl(prog.consts[idx]); break;
        case VK_SYS: b = "U[" + std::to_string(idx) + "]"; break;
        case VK_LINE0: b = "L0[" + std::to_string(idx) + " * l0s + e]"; break;
        case VK_LINE1: b = "L1[" + std::to_string(idx) + " * l1s]"; break;
        case VK_REG: b = "R[" + std::to_string(idx) + "][e]"; break;
        default: return std::string();
      }
      std::string x;
      switch (op) {
        case OP_LD: x = b; break;
        case OP_ADD: x = "__dadd_rn(A[e], " + b + ")"; break;
        case OP_SUB: x = "__dsub_rn(A[e], " + b + ")"; break;
        case OP_RSUB: x = "__dsub_rn(" + b + ", A[e])"; break;
        case OP_MUL: x = "__dmul_rn(A[e], " + b + ")"; break;
        case OP_DIV: x = "__ddiv_rn(A[e], " + b + ")"; break;
        case OP_RDIV: x = "__ddiv_rn(" + b + ", A[e])"; break;
        case OP_POW: x = "pow(A[e], " + b + ")"; break;
        case OP_RPOW: x = "pow(" + b + ", A[e])"; break;
        case OP_MIN: x = "((" + b + " < A[e]) ? " + b + " : A[e])"; break;
        case OP_MAX: x = "((A[e] < " + b + ") ? " + b + " : A[e])"; break;
        case OP_LT: x = "((A[e] < " + b + ") ? 1.0 : 0.0)"; break;
        case OP_GT: x = "((A[e] > " + b + ") ? 1.0 : 0.0)"; break;
        case OP_LE: x = "((A[e] <= " + b + ") ? 1.0 : 0.0)"; break;
        case OP_GE: x = "((A[e] >= " + b + ") ? 1.0 : 0.0)"; break;
        case OP_EQ: x = "((A[e] == " + b + ") ? 1.0 : 0.0)"; break;
        case OP_NE: x = "((A[e] != " + b + ") ? 1.0 : 0.0)"; break;
        case OP_LAND: x = "(((A[e] != 0.0) && (" + b + " != 0.0)) ? 1.0 : 0.0)"; break;
        case OP_LOR: x = "(((A[e] != 0.0) || (" + b + " != 0.0)) ? 1.0 : 0.0)"; break;
        case OP_RNDAND: x = "((vm_round(A[e]) && vm_round(" + b + ")) ? 1.0 : 0.0)"; break;
        case OP_RNDOR: x = "((vm_round(A[e]) || vm_round(" + b + ")) ? 1.0 : 0.0)"; break;
        case OP_ATAN2: x = "atan2(A[e], " + b + ")"; break;
        default: x = "atan2(" + b + ", A[e])"; break;  // OP_RATAN2
      }
      o << lanes << "A[e] = " << x << ";\n";
      continue;
    }
    if (op == OP_ST) {
      if (kind != VK_REG) return std::string();
      o << lanes << "R[" << idx << "][e] = A[e];\n";
      continue;
    }
    if (op == OP_SEL) {
      const uint32_t idx2 = (ins >> 21) & 0x3ffu;
      o << lanes << "if (vm_round(R[" << idx << "][e])) A[e] = R[" << idx2 << "][e];\n";
      continue;
    }
    std::string x;
    switch (op) {
      case OP_NEG: x = "-A[e]"; break;
      case OP_SQR: x = "__dmul_rn(A[e], A[e])"; break;
      case OP_CUBE: x = "__dmul_rn(__dmul_rn(A[e], A[e]), A[e])"; break;
      case OP_POW4: x = "__dmul_rn(__dmul_rn(__dmul_rn(A[e], A[e]), A[e]), A[e])"; break;
      case OP_SIGN: x = "((A[e] < 0.0) ? -1.0 : ((A[e] > 0.0) ? 1.0 : 0.0))"; break;
      case OP_RINT: x = "floor(__dadd_rn(A[e], 0.5))"; break;
      case OP_INT: x = "static_cast<double>(vm_round(A[e]))"; break;
      case OP_COT: x = "__ddiv_rn(1.0, tan(A[e]))"; break;
      case OP_CSC: x = "__ddiv_rn(1.0, sin(A[e]))"; break;
      case OP_SEC: x = "__ddiv_rn(1.0, cos(A[e]))"; break;
      default: {
        const char* fn = unary_fn(op);
        if (!fn) return std::string();
        x = std::string(fn) + "(A[e])";
      }
    }
    o << lanes << "A[e] = " << x << ";\n";
  }
  o << "}\n";
  return o.str();
}

// The kernel text mirrors body_cell_moments_vec (assemble_bodies.cuh) statement by statement; the
// GPU test suite checks the two bit for bit.
std::string cell_moments_source(const MomentPrograms& mp, int q, const double* qtab, const double* qsum,
                                bool* unrolled_rows) {
  const int mode = !mp.has_map ? 2 : (mp.jac_depends_on_y ? 0 : 1);  // MODE_GENERAL / MODE_CONST_F / MODE_NO_MAP
  const VmProgram& prog = mode == 0 ? *mp.full : *mp.G;
  const std::string fn_point = vm_to_cuda(prog, "prog_point", q);
  if (fn_point.empty()) return std::string();
  std::string fn_F;
  if (mode == 1) {
    fn_F = vm_to_cuda(*mp.F, "prog_F", 1);
    if (fn_F.empty()) return std::string();
  }
  // Rows of quadrature points are unrolled completely (table entries become immediate
  // constant-bank operands) unless the program is so long that the code would not fit the
  // instruction cache.
  const bool unroll_rows = static_cast<long long>(prog.n_code + 12) * q * q <= 12000;
  if (unrolled_rows) *unrolled_rows = unroll_rows;

  std::ostringstream o;
  o << "// generated by libmsgpu (jit.cpp): cell moments, q = " << q << ", mode = " << mode << "\n";
  o << "#define Q " << q << "\n#define MODE " << mode << "\n";
  o << "#define CELL_STRIDE " << CELL_STRIDE << "\n#define QTAB_STRIDE " << QTAB_STRIDE << "\n";
  o << "#define NREG_POINT " << (prog.n_regs < 1 ? 1 : prog.n_regs) << "\n#define OUT0_POINT " << prog.out0 << "\n";
  if (mode == 1)
    o << "#define NREG_F " << (mp.F->n_regs < 1 ? 1 : mp.F->n_regs) << "\n#define OUT0_F " << mp.F->out0 << "\n";
  o << R"(
struct MeshDims { int nc, m, n, ld, ncells, q, P; double h; };
struct Tables { const double* slots; const double* T0; const double* T1; int n_slots, n0, n1; };
struct PullBack { double det, k00, k01, k11; };

__device__ __forceinline__ int vm_round(double v) { return static_cast<int>(v + ((v >= 0.0) ? 0.5 : -0.5)); }

__device__ __forceinline__ PullBack pullback(double f00, double f01, double f10, double f11) {
  PullBack p;
  p.det = f00 * f11 - f10 * f01;
  const double inv_det = 1.0 / p.det;
  const double i00 = f11 * inv_det, i01 = -f01 * inv_det, i10 = -f10 * inv_det, i11 = f00 * inv_det;
  p.k00 = i00 * i00 + i01 * i01;
  p.k01 = i00 * i10 + i01 * i11;
  p.k11 = i10 * i10 + i11 * i11;
  return p;
}

__device__ __forceinline__ void moments_to_local(const double* mom, double scale, double* a) {
  const double A0 = mom[0], A1 = mom[1], A2 = mom[2];
  const double C0 = mom[3], C1 = mom[4], C2 = mom[5];
  const double B00 = mom[6], B01 = mom[7], B10 = mom[8], B11 = mom[9];
  a[0] = scale * (A0 + C0 + 2.0 * B00);
  a[1] = scale * (-A0 + C1 + B01 - B00);
  a[2] = scale * (A1 - C0 - B00 + B10);
  a[3] = scale * (-A1 - C1 - B01 - B10);
  a[4] = scale * (A0 + C2 - 2.0 * B01);
  a[5] = scale * (-A1 - C1 + B00 + B11);
  a[6] = scale * (A1 - C2 + B01 - B11);
  a[7] = scale * (A2 + C0 - 2.0 * B10);
  a[8] = scale * (-A2 + C1 - B11 + B10);
  a[9] = scale * (A2 + C2 + 2.0 * B11);
}
)";
  o << "__constant__ double c_qtab[" << q * q * QTAB_STRIDE << "] = {\n";
  for (int i = 0; i < q * q * QTAB_STRIDE; ++i) {
    char buf[48];
    std::snprintf(buf, sizeof buf, "%a,%s", qtab[i], (i % 7 == 6) ? "\n" : " ");
    o << buf;
  }
  o << "};\n__constant__ double c_qsum[" << QTAB_STRIDE << "] = {";
  for (int i = 0; i < QTAB_STRIDE; ++i) {
    char buf[48];
    std::snprintf(buf, sizeof buf, "%a, ", qsum[i]);
    o << buf;
  }
  o << "};\n#define QT(q, i) c_qtab[(q) * QTAB_STRIDE + (i)]\n";
  {
    // 1-D factors of the table entries (fe_tables.h build_q1d): moments are summed row by row
    std::vector<double> pt, wt;
    gauss_legendre_unit(q, pt, wt);
    const std::vector<double> q1d = build_q1d(q, pt, wt);
    o << "__constant__ double c_q1d[" << q1d.size() << "] = {";
    for (double v : q1d) {
      char buf[48];
      std::snprintf(buf, sizeof buf, "%a, ", v);
      o << buf;
    }
    o << "};\n#define QL(r, i) c_q1d[(r) * Q + (i)]\n\n";
  }
  o << fn_F << "\n" << fn_point << "\n";
  o << "#define UNROLL_ROWS " << (unroll_rows ? "_Pragma(\"unroll\")" : "_Pragma(\"unroll 1\")") << "\n";
  o << R"(
extern "C" __global__ void __launch_bounds__(128) k_cell_moments_jit(MeshDims md, Tables tb, int k0, int nk,
                                                                      double stiffness_scale,
                                                                      double* __restrict__ cell,
                                                                      int* __restrict__ error_flag) {
  const long long t = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (t >= static_cast<long long>(nk) * md.ncells) return;
  const int kl = static_cast<int>(t / md.ncells), c = static_cast<int>(t % md.ncells);
  const int cx = c / md.nc, cy = c % md.nc;
  const int k = k0 + kl;
  const double* U = tb.slots + static_cast<size_t>(k) * tb.n_slots;
  const double* T0k = tb.T0 + static_cast<size_t>(k) * tb.n0 * md.P + cx * Q;
  const double* T1k = tb.T1 + static_cast<size_t>(k) * tb.n1 * md.P + cy;

  double acc[CELL_STRIDE];
#pragma unroll
  for (int i = 0; i < CELL_STRIDE; ++i) acc[i] = 0.0;
  bool bad = false;
  double det_c = 1.0, m00_c = 1.0, m01_c = 0.0, m11_c = 1.0;
#if MODE == 1
  {
    double RF[NREG_F][1];
    prog_F(U, T0k, md.P, T1k, md.P, RF);
    const PullBack pb = pullback(RF[OUT0_F + 0][0], RF[OUT0_F + 1][0], RF[OUT0_F + 2][0], RF[OUT0_F + 3][0]);
    det_c = pb.det;
    m00_c = pb.k00 * det_c;
    m01_c = pb.k01 * det_c;
    m11_c = pb.k11 * det_c;
    bad |= !(det_c > 0.0);
  }
#endif
#if MODE != 0
  // F constant per system: only the load moments need the points.  phi_i(xi, eta) w = wa[i & 1](xi) wb[i >> 1](eta):
  // row sums first, the four products once per row, det(F) once per cell
  UNROLL_ROWS
  for (int qy = 0; qy < Q; ++qy) {
    double R[NREG_POINT][Q];
    prog_point(U, T0k, md.P, T1k + qy * md.nc, md.P, R);
    double r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int e = 0; e < Q; ++e) {
      const double g = R[OUT0_POINT][e];
      r0 += g * QL(0, e);
      r1 += g * QL(1, e);
    }
    const double b0 = QL(0, qy), b1 = QL(1, qy);
    acc[10] += r0 * b0;
    acc[11] += r1 * b0;
    acc[12] += r0 * b1;
    acc[13] += r1 * b1;
  }
#pragma unroll
  for (int i = 10; i < 14; ++i) acc[i] *= det_c;
#else
  // F varies over the cell.  Every table entry is a product of a factor in xi and a factor in eta (fe_tables.h):
  // row sums with the xi factors, the eta factors once per row - 11 instead of 19 FP64 operations per point
  // besides the pull-back itself
  UNROLL_ROWS
  for (int qy = 0; qy < Q; ++qy) {
    double R[NREG_POINT][Q];
    prog_point(U, T0k, md.P, T1k + qy * md.nc, md.P, R);
    double sa = 0.0, sc0 = 0.0, sc1 = 0.0, sc2 = 0.0, sb0 = 0.0, sb1 = 0.0, sd0 = 0.0, sd1 = 0.0, sg0 = 0.0, sg1 = 0.0;
#pragma unroll
    for (int e = 0; e < Q; ++e) {
      const PullBack pb = pullback(R[OUT0_POINT + 0][e], R[OUT0_POINT + 1][e], R[OUT0_POINT + 2][e], R[OUT0_POINT + 3][e]);
      const double g = R[OUT0_POINT + 4][e];
      const double det = pb.det;
      const double m00 = pb.k00 * det, m01 = pb.k01 * det, m11 = pb.k11 * det;
      bad |= !(det > 0.0);
      sa += m00 * QL(5, e);
      sc0 += m11 * QL(2, e);
      sc1 += m11 * QL(3, e);
      sc2 += m11 * QL(4, e);
      sb0 += m01 * QL(0, e);
      sb1 += m01 * QL(1, e);
      sd0 += det * QL(0, e);
      sd1 += det * QL(1, e);
      const double gd = g * det;
      sg0 += gd * QL(0, e);
      sg1 += gd * QL(1, e);
    }
    const double b0 = QL(0, qy), b1 = QL(1, qy);
    acc[0] += sa * QL(2, qy);
    acc[1] += sa * QL(3, qy);
    acc[2] += sa * QL(4, qy);
    acc[3] += sc0 * QL(5, qy);
    acc[4] += sc1 * QL(5, qy);
    acc[5] += sc2 * QL(5, qy);
    acc[6] += sb0 * b0;
    acc[7] += sb1 * b0;
    acc[8] += sb0 * b1;
    acc[9] += sb1 * b1;
    acc[10] += sg0 * b0;
    acc[11] += sg1 * b0;
    acc[12] += sg0 * b1;
    acc[13] += sg1 * b1;
    acc[14] += sd0 * b0;
    acc[15] += sd1 * b0;
    acc[16] += sd0 * b1;
    acc[17] += sd1 * b1;
  }
#endif
#if MODE != 0
#pragma unroll
  for (int i = 0; i < 3; ++i) acc[i] = m00_c * c_qsum[i];
#pragma unroll
  for (int i = 3; i < 6; ++i) acc[i] = m11_c * c_qsum[i];
#pragma unroll
  for (int i = 6; i < 10; ++i) acc[i] = m01_c * c_qsum[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[14 + i] = det_c * c_qsum[10 + i];
#endif
  double a[10];
  moments_to_local(acc, stiffness_scale, a);
  const double h2 = md.h * md.h;
  double* o = cell + static_cast<size_t>(kl) * CELL_STRIDE * md.ncells + c;
#if MODE != 0
  // F is constant per system: the matrix entries and the det-weighted basis integrals are the same in
  // every cell and the gather reads them from cell 0 (GatherArgs::uniform_stiffness)
  if (c == 0) {
#pragma unroll
    for (int e = 0; e < 10; ++e) o[static_cast<size_t>(e) * md.ncells] = a[e];
#pragma unroll
    for (int e = 14; e < CELL_STRIDE; ++e) o[static_cast<size_t>(e) * md.ncells] = acc[e] * h2;
  }
#pragma unroll
  for (int e = 10; e < 14; ++e) o[static_cast<size_t>(e) * md.ncells] = acc[e] * h2;
#else
#pragma unroll
  for (int e = 0; e < 10; ++e) o[static_cast<size_t>(e) * md.ncells] = a[e];
#pragma unroll
  for (int e = 10; e < CELL_STRIDE; ++e) o[static_cast<size_t>(e) * md.ncells] = acc[e] * h2;
#endif
  if (bad) *error_flag = 1;
}
)";
  return o.str();
}

// The kernel text mirrors body_cell_errors (extra_bodies.cuh); the five evaluations of the exact
// solution per quadrature point (value + central differences) run as five lanes of one program.
std::string cell_errors_source(const VmProgram& prog) {
  const std::string fn = vm_to_cuda(prog, "prog_exact", 5);
  if (fn.empty()) return std::string();
  std::ostringstream o;
  o << "// generated by libmsgpu (jit_codegen.cpp): per-cell error integrals\n";
  o << "#define NREG " << (prog.n_regs < 2 ? 2 : prog.n_regs) << "\n#define OUT0 " << prog.out0 << "\n";
  o << R"(
struct MeshDims { int nc, m, n, ld, ncells, q, P; double h; };
__device__ __forceinline__ int vm_round(double v) { return static_cast<int>(v + ((v >= 0.0) ? 0.5 : -0.5)); }
)";
  // the generated function declares R[n_regs][5]; make the declared size match NREG
  std::string f = fn;
  const std::string decl = "double (&R)[" + std::to_string(prog.n_regs < 1 ? 1 : prog.n_regs) + "][5]";
  const size_t pos = f.find(decl);
  if (pos == std::string::npos) return std::string();
  f.replace(pos, decl.size(), "double (&R)[NREG][5]");
  o << f << "\n";
  o << R"(
extern "C" __global__ void __launch_bounds__(128) k_cell_errors_jit(MeshDims md, const double* __restrict__ slots,
                                                                     int n_slots, int qe,
                                                                     const double* __restrict__ pt,
                                                                     const double* __restrict__ wt,
                                                                     const double* __restrict__ x, int N, double fd,
                                                                     double* __restrict__ out) {
  const long long gtid = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (gtid >= static_cast<long long>(N) * md.ncells) return;
  const int k = static_cast<int>(gtid / md.ncells), c = static_cast<int>(gtid % md.ncells);
  const int cx = c / md.nc, cy = c % md.nc;
  const double* U = slots + static_cast<size_t>(k) * n_slots;
  const double* xk = x + static_cast<size_t>(k) * md.n;
  const double v0 = xk[cx * md.m + cy], v1 = xk[(cx + 1) * md.m + cy];
  const double v2 = xk[cx * md.m + cy + 1], v3 = xk[(cx + 1) * md.m + cy + 1];
  const double ox = -1.0 + cx * md.h, oy = -1.0 + cy * md.h;
  double l2 = 0.0, semi = 0.0;
  for (int qy = 0; qy < qe; ++qy)
    for (int qx = 0; qx < qe; ++qx) {
      const double xi = pt[qx], eta = pt[qy];
      const double y0 = ox + md.h * xi, y1 = oy + md.h * eta;
      const double jxw = wt[qx] * wt[qy] * md.h * md.h;
      const double vh = v0 * ((1 - xi) * (1 - eta)) + v1 * (xi * (1 - eta)) + v2 * ((1 - xi) * eta) + v3 * (xi * eta);
      const double gx = (-(1 - eta) * v0 + (1 - eta) * v1 - eta * v2 + eta * v3) / md.h;
      const double gy = (-(1 - xi) * v0 - xi * v1 + (1 - xi) * v2 + xi * v3) / md.h;
      double R[NREG][5];
      R[0][0] = y0;      R[1][0] = y1;
      R[0][1] = y0 + fd; R[1][1] = y1;
      R[0][2] = y0 - fd; R[1][2] = y1;
      R[0][3] = y0;      R[1][3] = y1 + fd;
      R[0][4] = y0;      R[1][4] = y1 - fd;
      prog_exact(U, nullptr, 0, nullptr, 0, R);
      const double s = R[OUT0][0];
      const double sx = (R[OUT0][1] - R[OUT0][2]) / (2 * fd);
      const double sy = (R[OUT0][3] - R[OUT0][4]) / (2 * fd);
      l2 += (vh - s) * (vh - s) * jxw;
      semi += ((gx - sx) * (gx - sx) + (gy - sy) * (gy - sy)) * jxw;
    }
  out[(static_cast<size_t>(k) * 2 + 0) * md.ncells + c] = l2;
  out[(static_cast<size_t>(k) * 2 + 1) * md.ncells + c] = semi;
}
)";
  return o.str();
}

}  // namespace msgpu
