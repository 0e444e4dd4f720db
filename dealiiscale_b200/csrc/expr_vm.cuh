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

}  // namespace msgpu
