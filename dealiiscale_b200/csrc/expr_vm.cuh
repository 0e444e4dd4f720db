// Device expression VM (kernel K0, inlined into the assembly / face / error kernels).
//
// Replaces mu::Parser::Eval() as called by MultiscaleFunctionParser::mvalue / mmap /
// mtensor_value (reference: src/tools/multiscale_function_parser.cpp:338-430).
// Every opcode is one IEEE-754 double operation (no FMA contraction across
// opcodes), so the value of an expression does not depend on how the compiler
// split it between the system / line / point programs.
//
// The same source is compiled for the host by tests/native (MSGPU_HOSTDEV
// expands to nothing there) so that the exact interpreter can be checked
// against the oracle without a GPU; the product library only ever runs it on
// the device.
#pragma once
#include <cmath>

#include "vm_ops.h"

#ifdef __CUDACC__
#define MSGPU_HOSTDEV __host__ __device__ __forceinline__
#else
#define MSGPU_HOSTDEV inline
#endif

namespace msgpu {

struct VmIo {
  const double* U;   // system slots of this thread's system
  const double* L0;  // line-0 table entry of this thread's y0 point (stride l0_stride between nodes)
  const double* L1;
  double* R;         // per-thread register file (stride r_stride between registers)
  double* OUT;       // global outputs of table-building programs
  int l0_stride, l1_stride, r_stride, out_stride;
};

MSGPU_HOSTDEV int vm_round(double v) { return static_cast<int>(v + ((v >= 0.0) ? 0.5 : -0.5)); }

// Runs code[pc...] until OP_END; returns the pc after the END.
MSGPU_HOSTDEV int vm_run(const VmProgram& prog, int pc, const VmIo& io) {
  double A = 0.0;
  for (;;) {
    const uint32_t ins = prog.code[pc++];
    const uint32_t op = ins & 0xffu;
    if (op == OP_END) break;
    const uint32_t kind = (ins >> 8) & 7u;
    const uint32_t idx = (ins >> 11) & 0x3ffu;
    if (op <= OP_LAST_BINARY) {
      double b;
      switch (kind) {
        case VK_CONST: b = prog.consts[idx]; break;
        case VK_SYS: b = io.U[idx]; break;
        case VK_LINE0: b = io.L0[(long)idx * io.l0_stride]; break;
        case VK_LINE1: b = io.L1[(long)idx * io.l1_stride]; break;
        default: b = io.R[(long)idx * io.r_stride]; break;
      }
      switch (op) {
        case OP_LD: A = b; break;
        case OP_ADD: A = A + b; break;
        case OP_SUB: A = A - b; break;
        case OP_RSUB: A = b - A; break;
        case OP_MUL: A = A * b; break;
        case OP_DIV: A = A / b; break;
        case OP_RDIV: A = b / A; break;
        case OP_POW: A = pow(A, b); break;
        case OP_RPOW: A = pow(b, A); break;
        case OP_MIN: A = (b < A) ? b : A; break;   // std::min(A, b)
        case OP_MAX: A = (A < b) ? b : A; break;   // std::max(A, b)
        case OP_LT: A = (A < b) ? 1.0 : 0.0; break;
        case OP_GT: A = (A > b) ? 1.0 : 0.0; break;
        case OP_LE: A = (A <= b) ? 1.0 : 0.0; break;
        case OP_GE: A = (A >= b) ? 1.0 : 0.0; break;
        case OP_EQ: A = (A == b) ? 1.0 : 0.0; break;
        case OP_NE: A = (A != b) ? 1.0 : 0.0; break;
        case OP_LAND: A = ((A != 0.0) && (b != 0.0)) ? 1.0 : 0.0; break;
        case OP_LOR: A = ((A != 0.0) || (b != 0.0)) ? 1.0 : 0.0; break;
        case OP_RNDAND: A = (vm_round(A) && vm_round(b)) ? 1.0 : 0.0; break;
        case OP_RNDOR: A = (vm_round(A) || vm_round(b)) ? 1.0 : 0.0; break;
        case OP_ATAN2: A = atan2(A, b); break;
        default: A = atan2(b, A); break;  // OP_RATAN2
      }
    } else {
      switch (op) {
        case OP_NEG: A = -A; break;
        case OP_SQR: A = A * A; break;
        case OP_CUBE: A = A * A * A; break;
        case OP_POW4: A = A * A * A * A; break;
        case OP_SIN: A = sin(A); break;
        case OP_COS: A = cos(A); break;
        case OP_TAN: A = tan(A); break;
        case OP_ASIN: A = asin(A); break;
        case OP_ACOS: A = acos(A); break;
        case OP_ATAN: A = atan(A); break;
        case OP_SINH: A = sinh(A); break;
        case OP_COSH: A = cosh(A); break;
        case OP_TANH: A = tanh(A); break;
        case OP_ASINH: A = asinh(A); break;
        case OP_ACOSH: A = acosh(A); break;
        case OP_ATANH: A = atanh(A); break;
        case OP_LOG2: A = log2(A); break;
        case OP_LOG10: A = log10(A); break;
        case OP_LOG: A = log(A); break;
        case OP_EXP: A = exp(A); break;
        case OP_SQRT: A = sqrt(A); break;
        case OP_SIGN: A = (A < 0.0) ? -1.0 : ((A > 0.0) ? 1.0 : 0.0); break;
        case OP_RINT: A = floor(A + 0.5); break;
        case OP_ABS: A = fabs(A); break;
        case OP_INT: A = static_cast<double>(vm_round(A)); break;
        case OP_CEIL: A = ceil(A); break;
        case OP_FLOOR: A = floor(A); break;
        case OP_COT: A = 1.0 / tan(A); break;
        case OP_CSC: A = 1.0 / sin(A); break;
        case OP_SEC: A = 1.0 / cos(A); break;
        case OP_ERFC: A = erfc(A); break;
        case OP_ST:
          if (kind == VK_OUT) io.OUT[(long)idx * io.out_stride] = A;
          else io.R[(long)idx * io.r_stride] = A;
          break;
        default: {  // OP_SEL
          const uint32_t idx2 = (ins >> 21) & 0x3ffu;
          if (vm_round(io.R[(long)idx * io.r_stride])) A = io.R[(long)idx2 * io.r_stride];
          break;
        }
      }
    }
  }
  return pc;
}


// Out-of-line wrappers for the expensive functions: the vectorised VM calls them W times per
// opcode; keeping them out of line keeps the interpreter small enough for the instruction cache.
#ifdef __CUDACC__
#define MSGPU_OUTOFLINE static __host__ __device__ __noinline__
#else
#define MSGPU_OUTOFLINE inline
#endif
#define MSGPU_WRAP1(name, expr) MSGPU_OUTOFLINE double name(double x) { return (expr); }
MSGPU_WRAP1(f_sin, sin(x))
MSGPU_WRAP1(f_cos, cos(x))
MSGPU_WRAP1(f_tan, tan(x))
MSGPU_WRAP1(f_asin, asin(x))
MSGPU_WRAP1(f_acos, acos(x))
MSGPU_WRAP1(f_atan, atan(x))
MSGPU_WRAP1(f_sinh, sinh(x))
MSGPU_WRAP1(f_cosh, cosh(x))
MSGPU_WRAP1(f_tanh, tanh(x))
MSGPU_WRAP1(f_asinh, asinh(x))
MSGPU_WRAP1(f_acosh, acosh(x))
MSGPU_WRAP1(f_atanh, atanh(x))
MSGPU_WRAP1(f_log2, log2(x))
MSGPU_WRAP1(f_log10, log10(x))
MSGPU_WRAP1(f_log, log(x))
MSGPU_WRAP1(f_exp, exp(x))
MSGPU_WRAP1(f_erfc, erfc(x))
MSGPU_OUTOFLINE double f_pow(double a, double b) { return pow(a, b); }
MSGPU_OUTOFLINE double f_atan2(double a, double b) { return atan2(a, b); }
#undef MSGPU_WRAP1

// Vectorised variant: evaluates the program for W consecutive y0 points at once (one row of
// quadrature points of a cell).  Each bytecode instruction is fetched and decoded once per W
// points instead of once per point, which removes most of the interpretive overhead.
//   LINE0 operand j of lane e : io.L0[idx * l0_stride + e]     (the W points are consecutive in T0)
//   REG   register  j of lane e: io.R[(idx * W + e) * r_stride]
// Everything else (constants, system slots, LINE1 entries) is the same for the W lanes.
template <int W>
MSGPU_HOSTDEV int vm_run_vec(const VmProgram& prog, int pc, const VmIo& io) {
  double A[W];
#pragma unroll
  for (int e = 0; e < W; ++e) A[e] = 0.0;
  for (;;) {
    const uint32_t ins = prog.code[pc++];
    const uint32_t op = ins & 0xffu;
    if (op == OP_END) break;
    const uint32_t kind = (ins >> 8) & 7u;
    const uint32_t idx = (ins >> 11) & 0x3ffu;
    if (op <= OP_LAST_BINARY) {
      double b[W];
      switch (kind) {
        case VK_LINE0: {
          const double* p = io.L0 + (long)idx * io.l0_stride;
#pragma unroll
          for (int e = 0; e < W; ++e) b[e] = p[e];
          break;
        }
        case VK_REG: {
          const double* p = io.R + (long)idx * W * io.r_stride;
#pragma unroll
          for (int e = 0; e < W; ++e) b[e] = p[(long)e * io.r_stride];
          break;
        }
        default: {
          const double s = (kind == VK_CONST) ? prog.consts[idx]
                           : (kind == VK_SYS) ? io.U[idx]
                                              : io.L1[(long)idx * io.l1_stride];
#pragma unroll
          for (int e = 0; e < W; ++e) b[e] = s;
          break;
        }
      }
#define MSGPU_VEC(expr)                        \
  {                                            \
    _Pragma("unroll") for (int e = 0; e < W; ++e) A[e] = (expr); \
  }                                            \
  break
#define MSGPU_VEC_SLOW(expr) MSGPU_VEC(expr)
      switch (op) {
        case OP_LD: MSGPU_VEC(b[e]);
        case OP_ADD: MSGPU_VEC(A[e] + b[e]);
        case OP_SUB: MSGPU_VEC(A[e] - b[e]);
        case OP_RSUB: MSGPU_VEC(b[e] - A[e]);
        case OP_MUL: MSGPU_VEC(A[e] * b[e]);
        case OP_DIV: MSGPU_VEC_SLOW(A[e] / b[e]);
        case OP_RDIV: MSGPU_VEC_SLOW(b[e] / A[e]);
        case OP_POW: MSGPU_VEC_SLOW(f_pow(A[e], b[e]));
        case OP_RPOW: MSGPU_VEC_SLOW(f_pow(b[e], A[e]));
        case OP_MIN: MSGPU_VEC((b[e] < A[e]) ? b[e] : A[e]);
        case OP_MAX: MSGPU_VEC((A[e] < b[e]) ? b[e] : A[e]);
        case OP_LT: MSGPU_VEC((A[e] < b[e]) ? 1.0 : 0.0);
        case OP_GT: MSGPU_VEC((A[e] > b[e]) ? 1.0 : 0.0);
        case OP_LE: MSGPU_VEC((A[e] <= b[e]) ? 1.0 : 0.0);
        case OP_GE: MSGPU_VEC((A[e] >= b[e]) ? 1.0 : 0.0);
        case OP_EQ: MSGPU_VEC((A[e] == b[e]) ? 1.0 : 0.0);
        case OP_NE: MSGPU_VEC((A[e] != b[e]) ? 1.0 : 0.0);
        case OP_LAND: MSGPU_VEC(((A[e] != 0.0) && (b[e] != 0.0)) ? 1.0 : 0.0);
        case OP_LOR: MSGPU_VEC(((A[e] != 0.0) || (b[e] != 0.0)) ? 1.0 : 0.0);
        case OP_RNDAND: MSGPU_VEC((vm_round(A[e]) && vm_round(b[e])) ? 1.0 : 0.0);
        case OP_RNDOR: MSGPU_VEC((vm_round(A[e]) || vm_round(b[e])) ? 1.0 : 0.0);
        case OP_ATAN2: MSGPU_VEC_SLOW(f_atan2(A[e], b[e]));
        default: MSGPU_VEC_SLOW(f_atan2(b[e], A[e]));
      }
    } else {
      switch (op) {
        case OP_NEG: MSGPU_VEC(-A[e]);
        case OP_SQR: MSGPU_VEC(A[e] * A[e]);
        case OP_CUBE: MSGPU_VEC(A[e] * A[e] * A[e]);
        case OP_POW4: MSGPU_VEC(A[e] * A[e] * A[e] * A[e]);
        case OP_SIN: MSGPU_VEC_SLOW(f_sin(A[e]));
        case OP_COS: MSGPU_VEC_SLOW(f_cos(A[e]));
        case OP_TAN: MSGPU_VEC_SLOW(f_tan(A[e]));
        case OP_ASIN: MSGPU_VEC_SLOW(f_asin(A[e]));
        case OP_ACOS: MSGPU_VEC_SLOW(f_acos(A[e]));
        case OP_ATAN: MSGPU_VEC_SLOW(f_atan(A[e]));
        case OP_SINH: MSGPU_VEC_SLOW(f_sinh(A[e]));
        case OP_COSH: MSGPU_VEC_SLOW(f_cosh(A[e]));
        case OP_TANH: MSGPU_VEC_SLOW(f_tanh(A[e]));
        case OP_ASINH: MSGPU_VEC_SLOW(f_asinh(A[e]));
        case OP_ACOSH: MSGPU_VEC_SLOW(f_acosh(A[e]));
        case OP_ATANH: MSGPU_VEC_SLOW(f_atanh(A[e]));
        case OP_LOG2: MSGPU_VEC_SLOW(f_log2(A[e]));
        case OP_LOG10: MSGPU_VEC_SLOW(f_log10(A[e]));
        case OP_LOG: MSGPU_VEC_SLOW(f_log(A[e]));
        case OP_EXP: MSGPU_VEC_SLOW(f_exp(A[e]));
        case OP_SQRT: MSGPU_VEC(sqrt(A[e]));
        case OP_SIGN: MSGPU_VEC((A[e] < 0.0) ? -1.0 : ((A[e] > 0.0) ? 1.0 : 0.0));
        case OP_RINT: MSGPU_VEC(floor(A[e] + 0.5));
        case OP_ABS: MSGPU_VEC(fabs(A[e]));
        case OP_INT: MSGPU_VEC(static_cast<double>(vm_round(A[e])));
        case OP_CEIL: MSGPU_VEC(ceil(A[e]));
        case OP_FLOOR: MSGPU_VEC(floor(A[e]));
        case OP_COT: MSGPU_VEC_SLOW(1.0 / f_tan(A[e]));
        case OP_CSC: MSGPU_VEC_SLOW(1.0 / f_sin(A[e]));
        case OP_SEC: MSGPU_VEC_SLOW(1.0 / f_cos(A[e]));
        case OP_ERFC: MSGPU_VEC_SLOW(f_erfc(A[e]));
        case OP_ST: {
          double* p = io.R + (long)idx * W * io.r_stride;
#pragma unroll
          for (int e = 0; e < W; ++e) p[(long)e * io.r_stride] = A[e];
          break;
        }
        default: {  // OP_SEL
          const uint32_t idx2 = (ins >> 21) & 0x3ffu;
          const double* c = io.R + (long)idx * W * io.r_stride;
          const double* t = io.R + (long)idx2 * W * io.r_stride;
#pragma unroll
          for (int e = 0; e < W; ++e)
            if (vm_round(c[(long)e * io.r_stride])) A[e] = t[(long)e * io.r_stride];
          break;
        }
      }
    }
#undef MSGPU_VEC
#undef MSGPU_VEC_SLOW
  }
  return pc;
}

}  // namespace msgpu
