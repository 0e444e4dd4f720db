// See expr_compile.h.  Grammar notes (muParser 2.2.4 as configured by the reference,
// src/tools/multiscale_function_parser.cpp:209-233):
//   priorities  "||"=1 "&&"=2 "|"=1 "&"=2  comparisons=4  "+ -"=5  "* /"=6  "^"=7 (right assoc.)
//   the sign operators bind like "* /":  -x1^2 == -(x1^2),  -a*b == (-a)*b
//   blanks are insignificant, also between a function name and "(" (:287-310)
#include "expr_compile.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <sstream>

#include "expr_vm.cuh"

namespace msgpu {

// ---------------------------------------------------------------- folding
// Constants are folded by running the VM itself on the host, so that folding has
// exactly the semantics of the opcode it replaces.
static double fold(uint32_t op, double a, double b) {
  static VmProgram p;  // too large for the stack of small threads; single-threaded use
  p.n_code = 3;
  p.consts[0] = a;
  p.consts[1] = b;
  p.code[0] = vm_encode(OP_LD, VK_CONST, 0);
  p.code[1] = (op <= OP_LAST_BINARY) ? vm_encode(op, VK_CONST, 1) : vm_encode(op, 0, 0);
  p.code[2] = vm_encode(OP_ST, VK_REG, 0);
  p.code[3] = vm_encode(OP_END, 0, 0);
  double r = 0;
  VmIo io{};
  io.R = &r;
  io.r_stride = 1;
  vm_run(p, 0, io);
  return r;
}

// ---------------------------------------------------------------- graph
int ProgramSet::intern(const DagNode& n) {
  char key[128];
  uint64_t bits;
  std::memcpy(&bits, &n.value, 8);
  std::snprintf(key, sizeof key, "%d|%u|%llx|%d|%d|%d|%d", (int)n.type, n.op, (unsigned long long)bits, n.var, n.a,
                n.b, n.c);
  auto it = intern_.find(key);
  if (it != intern_.end()) return it->second;
  nodes_.push_back(n);
  int id = static_cast<int>(nodes_.size()) - 1;
  intern_.emplace(key, id);
  return id;
}
int ProgramSet::make_const(double v) {
  DagNode n{NT_CONST, 0, v, -1, -1, -1, -1, 0u};
  return intern(n);
}
int ProgramSet::make_var(int var) {
  unsigned dep = (var == VAR_X0 || var == VAR_X1) ? DEP_X : (var == VAR_T) ? DEP_T : (var == VAR_Y0) ? DEP_Y0 : DEP_Y1;
  DagNode n{NT_VAR, 0, 0.0, var, -1, -1, -1, dep};
  return intern(n);
}
int ProgramSet::make_unary(uint32_t op, int a) {
  if (nodes_[a].type == NT_CONST) return make_const(fold(op, nodes_[a].value, 0.0));
  DagNode n{NT_UNARY, op, 0.0, -1, a, -1, -1, nodes_[a].dep};
  return intern(n);
}
int ProgramSet::make_binary(uint32_t op, int a, int b) {
  if (nodes_[a].type == NT_CONST && nodes_[b].type == NT_CONST)
    return make_const(fold(op, nodes_[a].value, nodes_[b].value));
  DagNode n{NT_BINARY, op, 0.0, -1, a, b, -1, nodes_[a].dep | nodes_[b].dep};
  return intern(n);
}
int ProgramSet::make_select(int c, int a, int b) {
  if (nodes_[c].type == NT_CONST) return vm_round(nodes_[c].value) ? a : b;
  DagNode n{NT_SELECT, OP_SEL, 0.0, -1, c, a, b, nodes_[c].dep | nodes_[a].dep | nodes_[b].dep};
  return intern(n);
}

// ---------------------------------------------------------------- parser
class Parser {
 public:
  Parser(ProgramSet& ps, const std::string& text, const std::map<std::string, double>& constants, bool time_dep,
         int y0_node, int y1_node)
      : ps_(ps), s_(text), consts_(constants), time_dep_(time_dep), y0_(y0_node), y1_(y1_node) {}

  int parse() {
    int n = expr(0);
    skip();
    if (i_ != s_.size()) error("unexpected token");
    return n;
  }

 private:
  ProgramSet& ps_;
  const std::string& s_;
  const std::map<std::string, double>& consts_;
  bool time_dep_;
  int y0_, y1_;
  size_t i_ = 0;

  [[noreturn]] void error(const std::string& what) const {
    std::ostringstream o;
    o << "cannot parse expression <" << s_ << ">: " << what << " at character " << i_;
    throw CompileError(o.str());
  }
  void skip() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t' || s_[i_] == '\r' || s_[i_] == '\n')) ++i_;
  }
  bool eat(char c) {
    skip();
    if (i_ < s_.size() && s_[i_] == c) {
      ++i_;
      return true;
    }
    return false;
  }

  struct Infix {
    const char* tok;
    int prio;
    bool right;
    uint32_t op;
  };
  const Infix* peek_infix() {
    static const Infix table[] = {
        {"||", 1, false, OP_LOR}, {"&&", 2, false, OP_LAND}, {"<=", 4, false, OP_LE}, {">=", 4, false, OP_GE},
        {"!=", 4, false, OP_NE},  {"==", 4, false, OP_EQ},   {"|", 1, false, OP_RNDOR}, {"&", 2, false, OP_RNDAND},
        {"<", 4, false, OP_LT},   {">", 4, false, OP_GT},    {"+", 5, false, OP_ADD}, {"-", 5, false, OP_SUB},
        {"*", 6, false, OP_MUL},  {"/", 6, false, OP_DIV},   {"^", 7, true, OP_POW}};
    skip();
    for (const Infix& t : table) {
      size_t n = std::strlen(t.tok);
      if (s_.compare(i_, n, t.tok) == 0) return &t;
    }
    return nullptr;
  }

  int power(int base, int exponent) {
    const DagNode& e = ps_.nodes_[exponent];
    if (e.type == NT_CONST) {
      if (e.value == 2.0) return ps_.make_unary(OP_SQR, base);
      if (e.value == 3.0) return ps_.make_unary(OP_CUBE, base);
      if (e.value == 4.0) return ps_.make_unary(OP_POW4, base);
    }
    return ps_.make_binary(OP_POW, base, exponent);
  }

  int expr(int min_prio) {
    int lhs = prefix();
    for (;;) {
      const Infix* t = peek_infix();
      if (!t || t->prio < min_prio) return lhs;
      i_ += std::strlen(t->tok);
      int rhs = expr(t->right ? t->prio : t->prio + 1);
      lhs = (t->op == OP_POW) ? power(lhs, rhs) : ps_.make_binary(t->op, lhs, rhs);
    }
  }

  int prefix() {
    skip();
    if (i_ < s_.size() && (s_[i_] == '-' || s_[i_] == '+')) {
      char sign = s_[i_++];
      int operand = expr(7);  // only "^" binds tighter than a sign
      return sign == '-' ? ps_.make_unary(OP_NEG, operand) : operand;
    }
    return atom();
  }

  static bool ident_start(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
  static bool ident_char(char c) { return ident_start(c) || (c >= '0' && c <= '9'); }

  int atom() {
    skip();
    if (i_ >= s_.size()) error("unexpected end");
    char c = s_[i_];
    if (c == '(') {
      ++i_;
      int n = expr(0);
      if (!eat(')')) error("missing ')'");
      return n;
    }
    if ((c >= '0' && c <= '9') || c == '.') {
      const char* b = s_.c_str() + i_;
      char* e = nullptr;
      double v = std::strtod(b, &e);
      if (e == b) error("malformed number");
      i_ += static_cast<size_t>(e - b);
      return ps_.make_const(v);
    }
    if (!ident_start(c)) error(std::string("unexpected character '") + c + "'");
    size_t b = i_;
    while (i_ < s_.size() && ident_char(s_[i_])) ++i_;
    std::string name = s_.substr(b, i_ - b);
    skip();
    if (i_ < s_.size() && s_[i_] == '(' && function_known(name)) {
      ++i_;
      std::vector<int> args;
      if (!eat(')')) {
        for (;;) {
          args.push_back(expr(0));
          if (eat(',')) continue;
          if (eat(')')) break;
          error("missing ')' after the arguments of " + name);
        }
      }
      return call(name, args);
    }
    if (name == "x0") return ps_.make_var(VAR_X0);
    if (name == "x1") return ps_.make_var(VAR_X1);
    if (name == "y0") return y0_ >= 0 ? y0_ : ps_.make_var(VAR_Y0);
    if (name == "y1") return y1_ >= 0 ? y1_ : ps_.make_var(VAR_Y1);
    if (name == "t" && time_dep_) return ps_.make_var(VAR_T);
    auto it = consts_.find(name);
    if (it != consts_.end()) return ps_.make_const(it->second);
    if (name == "_pi") return ps_.make_const(3.141592653589793238462643);
    if (name == "_e") return ps_.make_const(2.718281828459045235360287);
    i_ = b;
    error("unknown name '" + name + "'");
  }

  static const std::map<std::string, uint32_t>& unary_functions() {
    static const std::map<std::string, uint32_t> t = {
        {"sin", OP_SIN},     {"cos", OP_COS},     {"tan", OP_TAN},     {"asin", OP_ASIN},   {"acos", OP_ACOS},
        {"atan", OP_ATAN},   {"sinh", OP_SINH},   {"cosh", OP_COSH},   {"tanh", OP_TANH},   {"asinh", OP_ASINH},
        {"acosh", OP_ACOSH}, {"atanh", OP_ATANH}, {"log2", OP_LOG2},   {"log10", OP_LOG10}, {"log", OP_LOG},
        {"ln", OP_LOG},      {"exp", OP_EXP},     {"sqrt", OP_SQRT},   {"sign", OP_SIGN},   {"rint", OP_RINT},
        {"abs", OP_ABS},     {"int", OP_INT},     {"ceil", OP_CEIL},   {"floor", OP_FLOOR}, {"cot", OP_COT},
        {"csc", OP_CSC},     {"sec", OP_SEC},     {"erfc", OP_ERFC}};
    return t;
  }
  static bool function_known(const std::string& n) {
    return unary_functions().count(n) || n == "min" || n == "max" || n == "sum" || n == "avg" || n == "if" ||
           n == "pow" || n == "atan2" || n == "rand" || n == "rand_seed";
  }
  int call(const std::string& name, const std::vector<int>& a) {
    auto u = unary_functions().find(name);
    if (u != unary_functions().end()) {
      if (a.size() != 1) error(name + "() takes one argument");
      return ps_.make_unary(u->second, a[0]);
    }
    if (name == "rand" || name == "rand_seed")
      error("rand()/rand_seed() are not reproducible and are not supported on the device");
    if (name == "pow" || name == "atan2") {
      if (a.size() != 2) error(name + "() takes two arguments");
      return ps_.make_binary(name == "pow" ? OP_POW : OP_ATAN2, a[0], a[1]);
    }
    if (name == "if") {
      if (a.size() != 3) error("if() takes three arguments");
      return ps_.make_select(a[0], a[1], a[2]);
    }
    if (a.empty()) error(name + "() needs at least one argument");
    uint32_t op = name == "min" ? OP_MIN : name == "max" ? OP_MAX : OP_ADD;
    int r = a[0];
    for (size_t k = 1; k < a.size(); ++k) r = ps_.make_binary(op, r, a[k]);
    if (name == "avg") r = ps_.make_binary(OP_DIV, r, ps_.make_const(static_cast<double>(a.size())));
    return r;
  }
};

std::vector<int> ProgramSet::parse_function(const std::string& text, const std::map<std::string, double>& constants,
                                            bool time_dependent, int n_components, int mapped_y0, int mapped_y1) {
  if (finalized_) throw CompileError("ProgramSet already finalized");
  // split on ';' — a trailing separator followed by blanks only is tolerated
  std::vector<std::string> parts;
  std::string cur;
  for (char ch : text) {
    if (ch == ';') {
      parts.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(ch);
    }
  }
  if (cur.find_first_not_of(" \t\r\n") != std::string::npos) parts.push_back(cur);
  if (static_cast<int>(parts.size()) != n_components) {
    std::ostringstream o;
    o << "expression <" << text << "> has " << parts.size() << " component(s), " << n_components << " expected";
    throw CompileError(o.str());
  }
  std::vector<int> out;
  for (const std::string& p : parts) {
    Parser ps(*this, p, constants, time_dependent, mapped_y0, mapped_y1);
    out.push_back(ps.parse());
  }
  return out;
}

int ProgramSet::add_point_program(const std::vector<int>& outputs, EvalMode mode) {
  if (finalized_) throw CompileError("ProgramSet already finalized");
  specs_.push_back({outputs, mode});
  return static_cast<int>(specs_.size()) - 1;
}

// ---------------------------------------------------------------- classification
ProgramSet::Cls ProgramSet::cls(int id, EvalMode mode) const {
  unsigned d = nodes_[id].dep;
  if (nodes_[id].type == NT_CONST) return CLS_CONST;
  if ((d & (DEP_Y0 | DEP_Y1)) == 0) return CLS_SYS;
  if (mode == MODE_POINTS) return CLS_POINT;
  if ((d & DEP_Y1) == 0) return CLS_LINE0;
  if ((d & DEP_Y0) == 0) return CLS_LINE1;
  return CLS_POINT;
}

// Walk the part of the DAG that the program of class `inline_cls` computes itself and record
// every node of another class it reads (those must be materialised in that class's table).
void ProgramSet::mark(int id, Cls inline_cls, EvalMode mode, std::vector<char>& seen) {
  if (id < 0) return;
  const DagNode& n = nodes_[id];
  if (n.type == NT_CONST) return;
  Cls c = cls(id, mode);
  bool is_input_var = n.type == NT_VAR && ((inline_cls == CLS_SYS) || (inline_cls == CLS_LINE0 && n.var == VAR_Y0) ||
                                           (inline_cls == CLS_LINE1 && n.var == VAR_Y1) ||
                                           (inline_cls == CLS_POINT && mode == MODE_POINTS &&
                                            (n.var == VAR_Y0 || n.var == VAR_Y1)));
  if (is_input_var) return;
  if (c != inline_cls) {
    std::map<int, int>& table = (c == CLS_SYS) ? sys_slot_ : (c == CLS_LINE0) ? l0_slot_ : l1_slot_;
    table.emplace(id, -1);
    return;
  }
  if (seen[id]) return;
  seen[id] = 1;
  mark(n.a, inline_cls, mode, seen);
  mark(n.b, inline_cls, mode, seen);
  mark(n.c, inline_cls, mode, seen);
}

// ---------------------------------------------------------------- code generation
class CodeGen {
 public:
  CodeGen(const ProgramSet& ps, ProgramSet::Cls inline_cls, EvalMode mode, VmProgram& out, int first_temp)
      : ps_(ps), icls_(inline_cls), mode_(mode), p_(out), next_temp_(first_temp), high_water_(first_temp) {
    p_.n_code = 0;
    n_consts_ = 0;
  }

  // Count how often each inline node is referenced, so shared ones are computed once.
  void count_uses(int id) {
    if (id < 0) return;
    if (!is_inline(id)) return;
    if (uses_[id]++ > 0) return;
    const DagNode& n = ps_.nodes_[id];
    count_uses(n.a);
    count_uses(n.b);
    count_uses(n.c);
  }
  void emit_store(int id, uint32_t kind, int index) {
    into_acc(id);
    emit(vm_encode(OP_ST, kind, index));
  }
  // After a table entry has been stored it can be read back through its table.
  void declare_stored(int id) { stored_.emplace(id, 1); }
  void finish() {
    emit(vm_encode(OP_END, 0, 0));
    p_.n_regs = high_water_;
    p_.reads_lines = mode_ == MODE_POINTS ? 1 : 0;  // MODE_POINTS: y0, y1 are per-thread inputs
    for (int pc = 0; pc < p_.n_code; ++pc) {
      const uint32_t op = p_.code[pc] & 0xffu, kind = (p_.code[pc] >> 8) & 7u;
      if (op != OP_END && op <= OP_LAST_BINARY && (kind == VK_LINE0 || kind == VK_LINE1)) p_.reads_lines = 1;
    }
  }

 private:
  const ProgramSet& ps_;
  ProgramSet::Cls icls_;
  EvalMode mode_;
  VmProgram& p_;
  int next_temp_, high_water_, n_consts_ = 0;
  std::vector<int> free_temps_;
  std::map<int, int> uses_;
  std::map<int, int> temp_of_;   // inline shared node -> register
  std::map<int, int> stored_;    // own-table nodes already written by this program

  void emit(uint32_t ins) {
    if (p_.n_code >= VM_MAX_CODE) throw CompileError("expression program too long for the device VM");
    p_.code[p_.n_code++] = ins;
  }
  int const_index(double v) {
    for (int i = 0; i < n_consts_; ++i)
      if (std::memcmp(&p_.consts[i], &v, 8) == 0) return i;
    if (n_consts_ >= VM_MAX_CONST) throw CompileError("too many constants for the device VM");
    p_.consts[n_consts_] = v;
    return n_consts_++;
  }
  int alloc_temp() {
    if (!free_temps_.empty()) {
      int t = free_temps_.back();
      free_temps_.pop_back();
      return t;
    }
    int t = next_temp_++;
    if (t >= 1023) throw CompileError("expression needs too many temporaries");
    high_water_ = std::max(high_water_, next_temp_);
    return t;
  }
  void free_temp(int t) { free_temps_.push_back(t); }

  bool input_var(const DagNode& n, int& reg) const {
    if (n.type != NT_VAR) return false;
    if (icls_ == ProgramSet::CLS_SYS) {
      reg = n.var == VAR_X0 ? VM_REG_IN0 : n.var == VAR_X1 ? VM_REG_IN1 : VM_REG_IN2;
      return n.var == VAR_X0 || n.var == VAR_X1 || n.var == VAR_T;
    }
    if (icls_ == ProgramSet::CLS_LINE0 && n.var == VAR_Y0) { reg = VM_REG_IN0; return true; }
    if (icls_ == ProgramSet::CLS_LINE1 && n.var == VAR_Y1) { reg = VM_REG_IN0; return true; }
    if (icls_ == ProgramSet::CLS_POINT && mode_ == MODE_POINTS) {
      if (n.var == VAR_Y0) { reg = VM_REG_IN0; return true; }
      if (n.var == VAR_Y1) { reg = VM_REG_IN1; return true; }
    }
    return false;
  }
  bool is_inline(int id) const {
    const DagNode& n = ps_.nodes_[id];
    if (n.type == NT_CONST) return false;
    int reg;
    if (input_var(n, reg)) return false;
    if (stored_.count(id)) return false;
    return ps_.cls(id, mode_) == icls_;
  }
  // If `id` can be used directly as an instruction operand, return its (kind, index).
  bool operand(int id, uint32_t& kind, uint32_t& index) {
    const DagNode& n = ps_.nodes_[id];
    if (n.type == NT_CONST) { kind = VK_CONST; index = const_index(n.value); return true; }
    int reg;
    if (input_var(n, reg)) { kind = VK_REG; index = reg; return true; }
    auto t = temp_of_.find(id);
    if (t != temp_of_.end()) { kind = VK_REG; index = t->second; return true; }
    ProgramSet::Cls c = ps_.cls(id, mode_);
    if (c != icls_ || stored_.count(id)) {
      const std::map<int, int>& table =
          (c == ProgramSet::CLS_SYS) ? ps_.sys_slot_ : (c == ProgramSet::CLS_LINE0) ? ps_.l0_slot_ : ps_.l1_slot_;
      auto it = table.find(id);
      if (it == table.end() || it->second < 0) throw CompileError("internal: node not materialised");
      kind = (c == ProgramSet::CLS_SYS) ? VK_SYS : (c == ProgramSet::CLS_LINE0) ? VK_LINE0 : VK_LINE1;
      index = it->second;
      return true;
    }
    return false;
  }
  static uint32_t reversed(uint32_t op) {
    switch (op) {
      case OP_ADD: case OP_MUL: case OP_MIN: case OP_MAX: case OP_EQ: case OP_NE:
      case OP_LAND: case OP_LOR: case OP_RNDAND: case OP_RNDOR: return op;
      case OP_SUB: return OP_RSUB;
      case OP_DIV: return OP_RDIV;
      case OP_POW: return OP_RPOW;
      case OP_ATAN2: return OP_RATAN2;
      case OP_LT: return OP_GT;
      case OP_GT: return OP_LT;
      case OP_LE: return OP_GE;
      case OP_GE: return OP_LE;
      default: throw CompileError("internal: operator cannot be reversed");
    }
  }

  // Generate code that leaves the value of node `id` in the accumulator.
  void into_acc(int id) {
    uint32_t kind, index;
    if (operand(id, kind, index)) {
      emit(vm_encode(OP_LD, kind, index));
      return;
    }
    const DagNode& n = ps_.nodes_[id];
    switch (n.type) {
      case NT_UNARY:
        into_acc(n.a);
        emit(vm_encode(n.op, 0, 0));
        break;
      case NT_BINARY: {
        uint32_t k2, i2;
        // NB: std::min/max are not symmetric for NaN / signed zeros only; "reversed" keeps them as they are.
        if (operand(n.b, k2, i2)) {
          into_acc(n.a);
          // operand() of n.b may have become a temporary while n.a was generated (shared node)
          operand(n.b, k2, i2);
          emit(vm_encode(n.op, k2, i2));
        } else if (operand(n.a, k2, i2)) {
          into_acc(n.b);
          operand(n.a, k2, i2);
          emit(vm_encode(reversed(n.op), k2, i2));
        } else {
          into_acc(n.b);
          uint32_t kb, ib;
          int t = -1;
          if (!operand(n.b, kb, ib)) {  // not already kept as a shared temporary
            t = alloc_temp();
            emit(vm_encode(OP_ST, VK_REG, t));
            kb = VK_REG;
            ib = t;
          }
          into_acc(n.a);
          emit(vm_encode(n.op, kb, ib));
          if (t >= 0) free_temp(t);
        }
        break;
      }
      case NT_SELECT: {
        int tc = alloc_temp();
        into_acc(n.a);
        emit(vm_encode(OP_ST, VK_REG, tc));
        int ta = alloc_temp();
        into_acc(n.b);
        emit(vm_encode(OP_ST, VK_REG, ta));
        into_acc(n.c);
        emit(vm_encode(OP_SEL, VK_REG, tc, ta));
        free_temp(ta);
        free_temp(tc);
        break;
      }
      default:
        throw CompileError("internal: unexpected node in code generation");
    }
    if (uses_[id] > 1) {  // shared inside this program: keep it
      int t = alloc_temp();
      emit(vm_encode(OP_ST, VK_REG, t));
      temp_of_[id] = t;  // never freed
    }
  }
};

void ProgramSet::finalize() {
  if (finalized_) return;
  // 1. what do the point programs read from the tables?
  for (const Spec& s : specs_) {
    std::vector<char> seen(nodes_.size(), 0);
    for (int o : s.outputs) mark(o, CLS_POINT, s.mode, seen);
  }
  // 2. what do the line programs read from the slot table?
  {
    std::vector<char> seen(nodes_.size(), 0);
    std::vector<int> roots;
    for (auto& kv : l0_slot_) roots.push_back(kv.first);
    for (int r : roots) {
      seen[r] = 0;
      const DagNode& n = nodes_[r];
      if (n.type == NT_VAR) continue;
      seen[r] = 1;
      mark(n.a, CLS_LINE0, MODE_LINES, seen);
      mark(n.b, CLS_LINE0, MODE_LINES, seen);
      mark(n.c, CLS_LINE0, MODE_LINES, seen);
    }
  }
  {
    std::vector<char> seen(nodes_.size(), 0);
    std::vector<int> roots;
    for (auto& kv : l1_slot_) roots.push_back(kv.first);
    for (int r : roots) {
      const DagNode& n = nodes_[r];
      if (n.type == NT_VAR) continue;
      seen[r] = 1;
      mark(n.a, CLS_LINE1, MODE_LINES, seen);
      mark(n.b, CLS_LINE1, MODE_LINES, seen);
      mark(n.c, CLS_LINE1, MODE_LINES, seen);
    }
  }
  // 3. table indices in node order (children precede parents by construction)
  auto assign = [](std::map<int, int>& table, std::vector<int>& order) {
    order.clear();
    for (auto& kv : table) {
      kv.second = static_cast<int>(order.size());
      order.push_back(kv.first);
    }
  };
  assign(sys_slot_, sys_nodes_);
  assign(l0_slot_, l0_nodes_);
  assign(l1_slot_, l1_nodes_);
  if (sys_nodes_.size() > 1000 || l0_nodes_.size() > 1000 || l1_nodes_.size() > 1000)
    throw CompileError("expression set needs too many table entries");
  slots_depend_on_t_ = false;
  for (int id : sys_nodes_)
    if (nodes_[id].dep & DEP_T) slots_depend_on_t_ = true;

  // 4. table-building programs
  auto build_table_program = [&](Cls c, const std::vector<int>& order, VmProgram& prog) {
    CodeGen g(*this, c, MODE_LINES, prog, VM_REG_TMP0);
    prog.out0 = 0;
    for (int id : order) g.count_uses(id);
    for (size_t i = 0; i < order.size(); ++i) {
      g.emit_store(order[i], VK_OUT, static_cast<int>(i));
      g.declare_stored(order[i]);
    }
    g.finish();
  };
  build_table_program(CLS_SYS, sys_nodes_, slot_prog_);
  build_table_program(CLS_LINE0, l0_nodes_, line0_prog_);
  build_table_program(CLS_LINE1, l1_nodes_, line1_prog_);

  // 5. point programs
  point_progs_.assign(specs_.size(), VmProgram{});
  for (size_t s = 0; s < specs_.size(); ++s) {
    const Spec& sp = specs_[s];
    // MODE_LINES programs have no register inputs; MODE_POINTS programs read y0, y1 from R[0], R[1]
    const int out0 = sp.mode == MODE_LINES ? 0 : 2;
    point_progs_[s].out0 = out0;
    CodeGen g(*this, CLS_POINT, sp.mode, point_progs_[s], out0 + static_cast<int>(sp.outputs.size()));
    for (int o : sp.outputs) g.count_uses(o);
    for (size_t i = 0; i < sp.outputs.size(); ++i) g.emit_store(sp.outputs[i], VK_REG, out0 + static_cast<int>(i));
    g.finish();
  }
  finalized_ = true;
}

std::string ProgramSet::disassemble(const VmProgram& p) const {
  static const char* names[] = {"END",  "LD",   "ADD",  "SUB",   "RSUB",  "MUL",   "DIV",  "RDIV",  "POW",   "RPOW",
                                "MIN",  "MAX",  "LT",   "GT",    "LE",    "GE",    "EQ",   "NE",    "LAND",  "LOR",
                                "RNDAND", "RNDOR", "ATAN2", "RATAN2", "NEG", "SQR", "CUBE", "POW4", "SIN",   "COS",
                                "TAN",  "ASIN", "ACOS", "ATAN",  "SINH",  "COSH",  "TANH", "ASINH", "ACOSH", "ATANH",
                                "LOG2", "LOG10", "LOG", "EXP",   "SQRT",  "SIGN",  "RINT", "ABS",   "INT",   "CEIL",
                                "FLOOR", "COT", "CSC",  "SEC",   "ERFC",  "ST",    "SEL"};
  static const char* kinds[] = {"c", "U", "L0", "L1", "R", "OUT", "?", "?"};
  std::ostringstream o;
  for (int i = 0; i < p.n_code; ++i) {
    uint32_t ins = p.code[i], op = ins & 0xff, kind = (ins >> 8) & 7, idx = (ins >> 11) & 0x3ff, idx2 = (ins >> 21) & 0x3ff;
    o << i << ": " << (op < OP_COUNT ? names[op] : "??");
    if (op != OP_END && (op <= OP_LAST_BINARY || op == OP_ST || op == OP_SEL)) {
      o << " " << kinds[kind] << "[" << idx << "]";
      if (kind == VK_CONST) o << "=" << p.consts[idx];
      if (op == OP_SEL) o << ", R[" << idx2 << "]";
    }
    o << "\n";
  }
  return o.str();
}

}  // namespace msgpu
