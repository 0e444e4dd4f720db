// Host-side expression compiler: .prm text -> bytecode for the device VM.
//
// Takes the role of MultiscaleFunctionParser::initialize / init_muparser
// (reference: src/tools/multiscale_function_parser.cpp:54-107, :198-324, :327-335):
// same grammar (muParser 2.2.4 operators and priorities, the deal.II compatibility
// functions, ';'-separated components, optional mapper composition), but instead
// of one byte-code per component it builds ONE expression DAG per context and
// splits every expression by what it depends on:
//
//   depends on            evaluated                         stored in
//   -------------------   -------------------------------   ---------------------------
//   nothing               at compile time (folded)          program constant table
//   x0,x1,t only          once per micro system             slot table   U [k][slot]
//   y0 (+x,t)             once per (system, 1-D y0 point)   line table   T0[k][node][p]
//   y1 (+x,t)             once per (system, 1-D y1 point)   line table   T1[k][node][p]
//   y0 and y1             at every quadrature point         per-thread registers
//
// On a tensor-product quadrature of a uniform square mesh there are only
// cells_per_axis*q distinct y0 (and y1) values, so the y0-only / y1-only parts of
// the mapping Jacobian are evaluated (cells_per_axis*q) times instead of
// (cells_per_axis*q)^2 times per system.  The split never changes the value of an
// expression: every node is the same IEEE double operation wherever it runs.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "vm_ops.h"

namespace msgpu {

struct CompileError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

enum DepBits : unsigned { DEP_X = 1u, DEP_T = 2u, DEP_Y0 = 4u, DEP_Y1 = 8u };
enum VarId : int { VAR_X0 = 0, VAR_X1 = 1, VAR_Y0 = 2, VAR_Y1 = 3, VAR_T = 4 };
enum NodeType : int { NT_CONST, NT_VAR, NT_UNARY, NT_BINARY, NT_SELECT };
enum EvalMode : int {
  MODE_LINES = 0,   // y0 / y1 come from the line tables (quadrature points and mesh vertices)
  MODE_POINTS = 1   // y0 / y1 are per-thread inputs in R[0], R[1] (arbitrary points, e.g. finite differences)
};

struct DagNode {
  NodeType type;
  uint32_t op;   // VmOp for unary / binary nodes (never an "R" variant)
  double value;  // NT_CONST
  int var;       // NT_VAR
  int a, b, c;   // children (node ids), -1 if unused
  unsigned dep;
};


class ProgramSet {
 public:
  // Parses `text` (one or more ';'-separated components).  If mapper components are given,
  // y0,y1 in `text` denote the MAPPED point, i.e. the function is composed with the mapping
  // (reference: MultiscaleFunctionParser::mvalue with a mapper, :342-347).
  // Returns the node id of every component.
  std::vector<int> parse_function(const std::string& text, const std::map<std::string, double>& constants,
                                  bool time_dependent, int n_components, int mapped_y0 = -1, int mapped_y1 = -1);

  // A point program evaluates `outputs` (node ids) into registers prog.out0, prog.out0+1, .. in order.
  int add_point_program(const std::vector<int>& outputs, EvalMode mode);

  // Decides what is materialised where and generates all programs.
  void finalize();

  int n_slots() const { return static_cast<int>(sys_nodes_.size()); }
  int n_line0() const { return static_cast<int>(l0_nodes_.size()); }
  int n_line1() const { return static_cast<int>(l1_nodes_.size()); }
  bool time_dependent_slots() const { return slots_depend_on_t_; }
  const VmProgram& slot_program() const { return slot_prog_; }
  const VmProgram& line0_program() const { return line0_prog_; }
  const VmProgram& line1_program() const { return line1_prog_; }
  const VmProgram& point_program(int id) const { return point_progs_.at(id); }
  int point_program_outputs(int id) const { return static_cast<int>(specs_.at(id).outputs.size()); }

  // Does node `id` depend on y at all / on x at all?  (used to detect shared matrices)
  unsigned node_dep(int id) const { return nodes_.at(id).dep; }
  std::string disassemble(const VmProgram& p) const;

 private:
  struct Spec {
    std::vector<int> outputs;
    EvalMode mode;
  };
  enum Cls { CLS_CONST, CLS_SYS, CLS_LINE0, CLS_LINE1, CLS_POINT };

  std::vector<DagNode> nodes_;
  std::unordered_map<std::string, int> intern_;
  std::vector<Spec> specs_;
  std::vector<int> sys_nodes_, l0_nodes_, l1_nodes_;
  std::map<int, int> sys_slot_, l0_slot_, l1_slot_;
  bool slots_depend_on_t_ = false;
  bool finalized_ = false;
  VmProgram slot_prog_{}, line0_prog_{}, line1_prog_{};
  std::vector<VmProgram> point_progs_;

  // graph construction (hash-consed, constants folded)
  int make_const(double v);
  int make_var(int var);
  int make_unary(uint32_t op, int a);
  int make_binary(uint32_t op, int a, int b);
  int make_select(int c, int a, int b);
  int intern(const DagNode& n);

  Cls cls(int id, EvalMode mode) const;
  void mark(int id, Cls inline_cls, EvalMode mode, std::vector<char>& seen);

  friend class Parser;
  friend class CodeGen;
};

}  // namespace msgpu
