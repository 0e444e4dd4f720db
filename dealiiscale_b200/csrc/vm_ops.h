// Bytecode of the device expression VM (kernel K0 of SURVEY.md §2c).
//
// Replaces the muParser byte-code evaluator the reference calls at every
// quadrature point (reference: src/tools/multiscale_function_parser.cpp:362, :401, :427).
//
// The machine has one accumulator A.  An instruction is 32 bits:
//     bits  0..7   opcode
//     bits  8..10  operand kind (where operand b comes from)
//     bits 11..20  operand index
//     bits 21..30  second index (OP_SEL only)
// Binary opcodes compute A = A (op) b, "R" variants compute A = b (op) A.
#pragma once
#include <cstdint>

namespace msgpu {

enum VmKind : uint32_t {
  VK_CONST = 0,  // program constant table
  VK_SYS = 1,    // per-system slot table        U[idx]
  VK_LINE0 = 2,  // per-(system, y0 point) table L0[idx * l0_stride]
  VK_LINE1 = 3,  // per-(system, y1 point) table L1[idx * l1_stride]
  VK_REG = 4,    // per-thread register file     R[idx * r_stride]
  VK_OUT = 5     // OP_ST only: global output    OUT[idx * out_stride]
};

enum VmOp : uint32_t {
  OP_END = 0,
  // ---- operand-taking opcodes ----
  OP_LD, OP_ADD, OP_SUB, OP_RSUB, OP_MUL, OP_DIV, OP_RDIV, OP_POW, OP_RPOW,
  OP_MIN, OP_MAX, OP_LT, OP_GT, OP_LE, OP_GE, OP_EQ, OP_NE,
  OP_LAND, OP_LOR,      // "&&" "||"  : operands compared with 0
  OP_RNDAND, OP_RNDOR,  // "&"  "|"   : deal.II's mu_and / mu_or on rounded operands
  OP_ATAN2, OP_RATAN2,
  OP_LAST_BINARY = OP_RATAN2,
  // ---- accumulator-only opcodes ----
  OP_NEG, OP_SQR, OP_CUBE, OP_POW4,
  OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_SINH, OP_COSH, OP_TANH,
  OP_ASINH, OP_ACOSH, OP_ATANH, OP_LOG2, OP_LOG10, OP_LOG, OP_EXP, OP_SQRT,
  OP_SIGN, OP_RINT, OP_ABS, OP_INT, OP_CEIL, OP_FLOOR, OP_COT, OP_CSC, OP_SEC, OP_ERFC,
  // ---- stores / select ----
  OP_ST,   // kind VK_REG: R[idx] = A ; kind VK_OUT: OUT[idx] = A
  OP_SEL,  // A = round(R[idx]) ? R[idx2] : A
  OP_COUNT
};

constexpr int VM_MAX_CODE = 1024;
constexpr int VM_MAX_CONST = 192;

// One compiled program, passed to kernels by value (const __grid_constant__):
// instruction fetches and constant operands are then uniform constant-bank loads.
struct VmProgram {
  int n_code;
  int n_regs;  // per-thread registers the program touches (inputs + outputs + temporaries)
  int out0;    // first output register of a point program (0 in MODE_LINES, 2 in MODE_POINTS)
  int reads_lines;  // 0: no operand comes from a line table - the outputs of a point program are the same at every point
  uint32_t code[VM_MAX_CODE];
  double consts[VM_MAX_CONST];
};

// Registers every program reserves for inputs.
constexpr int VM_REG_IN0 = 0;  // x0 (slot program) | y0 (line-0 program, MODE_POINTS) | y1 (line-1 program)
constexpr int VM_REG_IN1 = 1;  // x1 (slot program) | y1 (MODE_POINTS)
constexpr int VM_REG_IN2 = 2;  // t  (slot program)
constexpr int VM_REG_TMP0 = 3; // first temporary of the table-building programs

inline constexpr uint32_t vm_encode(uint32_t op, uint32_t kind, uint32_t idx, uint32_t idx2 = 0) {
  return op | (kind << 8) | (idx << 11) | (idx2 << 21);
}

}  // namespace msgpu
