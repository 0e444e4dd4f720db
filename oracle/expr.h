// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Tree-walking evaluator for the text expressions found in the reference's
// .prm files.  It restates the observable behaviour of muParser 2.2.4 as it is
// configured by the reference (reference: src/tools/multiscale_function_parser.cpp:198-324
// for the set-up, :338-430 for mvalue / mmap / mtensor_value).  muParser itself
// is a third-party dependency that is NOT vendored under /root/reference
// (deal.II 9.1.1 bundles muParser 2.2.4, CMakeLists.txt:13); what is restated
// here is its published grammar:
//   * binary operators and priorities: "||"(1) "&&"(2) "|"(1, deal.II) "&"(2, deal.II)
//     "<= >= != == < >"(4)  "+ -"(5)  "* /"(6)  "^"(7, right associative)
//   * sign operators "-" "+" bind like "* /" (prINFIX == prMUL_DIV), i.e.
//     "-x1^2" == "-(x1^2)"
//   * functions: sin cos tan asin acos atan sinh cosh tanh asinh acosh atanh
//     log2 log10 log ln exp sqrt sign rint abs min max sum avg  (muParser)
//     if int ceil cot csc floor sec pow erfc                    (deal.II, file:line above :220-233)
//     rand / rand_seed are refused (non-deterministic).
//   * constants "_pi" "_e" plus whatever the caller injects.
//   * components separated by ';' (trailing ';' tolerated), blanks ignored.
// Pinned by the reference's own parser tests (tests/test_tools.cpp:21-40), see
// tests/test_oracle_expr.py.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct ExprError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

enum class NodeKind { Const, Var, Unary, Binary, Call };

struct Node {
  NodeKind kind;
  double value = 0;      // Const
  int var = -1;          // Var: index into the variable vector
  std::string op;        // Unary / Binary / Call
  int code = -1;         // op resolved to an index into op_names() at parse time
  std::vector<std::unique_ptr<Node>> args;
};

inline int mu_round(double v) { return static_cast<int>(v + ((v >= 0.0) ? 0.5 : -0.5)); }

class Expression {
 public:
  Expression() = default;
  Expression(const std::string& text, const std::vector<std::string>& vars,
             const std::map<std::string, double>& constants) {
    parse(text, vars, constants);
  }

  void parse(const std::string& text, const std::vector<std::string>& vars,
             const std::map<std::string, double>& constants) {
    src_ = text;
    vars_ = vars;
    consts_ = constants;
    consts_.emplace("_pi", 3.141592653589793238462643);
    consts_.emplace("_e", 2.718281828459045235360287);
    pos_ = 0;
    root_ = parse_level(0);
    skip_ws();
    if (pos_ != src_.size()) fail("unexpected token");
    resolve(*root_);
  }

  double eval(const double* v) const { return eval_node(*root_, v); }
  bool empty() const { return !root_; }
  const std::string& source() const { return src_; }

 private:
  std::string src_;
  std::vector<std::string> vars_;
  std::map<std::string, double> consts_;
  size_t pos_ = 0;
  std::unique_ptr<Node> root_;

  [[noreturn]] void fail(const std::string& msg) const {
    throw ExprError("expression parse error: " + msg + " at position " + std::to_string(pos_) +
                    " in <" + src_ + ">");
  }
  void skip_ws() {
    while (pos_ < src_.size() && (src_[pos_] == ' ' || src_[pos_] == '\t')) ++pos_;
  }
  bool match(const char* s) {
    skip_ws();
    size_t n = std::char_traits<char>::length(s);
    if (src_.compare(pos_, n, s) == 0) {
      pos_ += n;
      return true;
    }
    return false;
  }

  // Binary operator table: text, priority, right-assoc.
  struct BinOp {
    const char* text;
    int prio;
    bool right;
  };
  static const std::vector<BinOp>& binops() {
    static const std::vector<BinOp> t = {
        {"||", 1, false}, {"&&", 2, false}, {"<=", 4, false}, {">=", 4, false}, {"!=", 4, false},
        {"==", 4, false}, {"|", 1, false},  {"&", 2, false},  {"<", 4, false},  {">", 4, false},
        {"+", 5, false},  {"-", 5, false},  {"*", 6, false},  {"/", 6, false},  {"^", 7, true}};
    return t;
  }
  const BinOp* peek_binop() {
    skip_ws();
    for (const auto& b : binops()) {
      size_t n = std::char_traits<char>::length(b.text);
      if (src_.compare(pos_, n, b.text) == 0) return &b;
    }
    return nullptr;
  }

  // Precedence climbing.  parse_level(p) parses an expression whose binary
  // operators all have priority >= p.
  std::unique_ptr<Node> parse_level(int min_prio) {
    auto lhs = parse_unary();
    for (;;) {
      const BinOp* b = peek_binop();
      if (!b || b->prio < min_prio) break;
      pos_ += std::char_traits<char>::length(b->text);
      auto rhs = parse_level(b->right ? b->prio : b->prio + 1);
      auto n = std::make_unique<Node>();
      n->kind = NodeKind::Binary;
      n->op = b->text;
      n->args.push_back(std::move(lhs));
      n->args.push_back(std::move(rhs));
      lhs = std::move(n);
    }
    return lhs;
  }

  // Sign operators have the priority of "* /": the operand of a sign is an
  // expression made of operators that bind tighter, i.e. only "^".
  std::unique_ptr<Node> parse_unary() {
    skip_ws();
    if (pos_ < src_.size() && (src_[pos_] == '-' || src_[pos_] == '+')) {
      char c = src_[pos_++];
      auto operand = parse_level(7);
      if (c == '+') return operand;
      auto n = std::make_unique<Node>();
      n->kind = NodeKind::Unary;
      n->op = "neg";
      n->args.push_back(std::move(operand));
      return n;
    }
    return parse_primary();
  }

  std::unique_ptr<Node> parse_primary() {
    skip_ws();
    if (pos_ >= src_.size()) fail("unexpected end of expression");
    char c = src_[pos_];
    if (c == '(') {
      ++pos_;
      auto e = parse_level(0);
      if (!match(")")) fail("missing ')'");
      return e;
    }
    if ((c >= '0' && c <= '9') || c == '.') {
      size_t used = 0;
      double v = 0;
      try {
        v = std::stod(src_.substr(pos_), &used);
      } catch (...) {
        fail("bad number");
      }
      pos_ += used;
      auto n = std::make_unique<Node>();
      n->kind = NodeKind::Const;
      n->value = v;
      return n;
    }
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_') {
      size_t b = pos_;
      while (pos_ < src_.size() &&
             ((src_[pos_] >= 'a' && src_[pos_] <= 'z') || (src_[pos_] >= 'A' && src_[pos_] <= 'Z') ||
              (src_[pos_] >= '0' && src_[pos_] <= '9') || src_[pos_] == '_'))
        ++pos_;
      std::string name = src_.substr(b, pos_ - b);
      skip_ws();
      if (pos_ < src_.size() && src_[pos_] == '(' && is_function(name)) {
        ++pos_;
        auto n = std::make_unique<Node>();
        n->kind = NodeKind::Call;
        n->op = name;
        skip_ws();
        if (pos_ < src_.size() && src_[pos_] == ')') {
          ++pos_;
        } else {
          for (;;) {
            n->args.push_back(parse_level(0));
            if (match(",")) continue;
            if (match(")")) break;
            fail("missing ')' in call of " + name);
          }
        }
        check_arity(*n);
        return n;
      }
      for (size_t i = 0; i < vars_.size(); ++i)
        if (vars_[i] == name) {
          auto n = std::make_unique<Node>();
          n->kind = NodeKind::Var;
          n->var = static_cast<int>(i);
          return n;
        }
      auto it = consts_.find(name);
      if (it != consts_.end()) {
        auto n = std::make_unique<Node>();
        n->kind = NodeKind::Const;
        n->value = it->second;
        return n;
      }
      pos_ = b;
      fail("unknown identifier '" + name + "'");
    }
    fail(std::string("unexpected character '") + c + "'");
  }

  static int fixed_arity(const std::string& f) {
    static const std::map<std::string, int> t = {
        {"sin", 1},   {"cos", 1},   {"tan", 1},   {"asin", 1},  {"acos", 1},  {"atan", 1},
        {"sinh", 1},  {"cosh", 1},  {"tanh", 1},  {"asinh", 1}, {"acosh", 1}, {"atanh", 1},
        {"log2", 1},  {"log10", 1}, {"log", 1},   {"ln", 1},    {"exp", 1},   {"sqrt", 1},
        {"sign", 1},  {"rint", 1},  {"abs", 1},   {"min", -1},  {"max", -1},  {"sum", -1},
        {"avg", -1},  {"if", 3},    {"int", 1},   {"ceil", 1},  {"cot", 1},   {"csc", 1},
        {"floor", 1}, {"sec", 1},   {"pow", 2},   {"erfc", 1},  {"atan2", 2}};
    auto it = t.find(f);
    return it == t.end() ? -2 : it->second;
  }
  static bool is_function(const std::string& f) {
    return fixed_arity(f) != -2 || f == "rand" || f == "rand_seed";
  }
  void check_arity(const Node& n) const {
    if (n.op == "rand" || n.op == "rand_seed")
      fail("rand()/rand_seed() are not reproducible and are refused");
    int a = fixed_arity(n.op);
    if (a == -1) {
      if (n.args.empty()) fail("function " + n.op + " needs at least one argument");
    } else if (static_cast<int>(n.args.size()) != a) {
      fail("wrong number of arguments for " + n.op);
    }
  }

  static const std::vector<std::string>& op_names() {
    static const std::vector<std::string> t = {
        "+", "-", "*", "/", "^", "<", ">", "<=", ">=", "==", "!=", "&&", "||", "&", "|",              // 0..14
        "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh",         // 15..25
        "atanh", "log2", "log10", "log", "ln", "exp", "sqrt", "sign", "rint", "abs", "int", "ceil",    // 26..37
        "floor", "cot", "csc", "sec", "erfc", "pow", "atan2", "if", "min", "max", "sum", "avg"};       // 38..49
    return t;
  }
  static int op_code(const std::string& o) {
    const auto& t = op_names();
    for (size_t i = 0; i < t.size(); ++i)
      if (t[i] == o) return static_cast<int>(i);
    throw ExprError("internal: unknown operator " + o);
  }
  static void resolve(Node& n) {
    if (n.kind == NodeKind::Binary || n.kind == NodeKind::Call) n.code = op_code(n.op);
    for (auto& a : n.args) resolve(*a);
  }

  static double eval_node(const Node& n, const double* v) {
    switch (n.kind) {
      case NodeKind::Const:
        return n.value;
      case NodeKind::Var:
        return v[n.var];
      case NodeKind::Unary:
        return -eval_node(*n.args[0], v);
      case NodeKind::Binary: {
        double a = eval_node(*n.args[0], v), b = eval_node(*n.args[1], v);
        switch (n.code) {
          case 0: return a + b;
          case 1: return a - b;
          case 2: return a * b;
          case 3: return a / b;
          case 4: return pow_like_muparser(a, b);
          case 5: return a < b;
          case 6: return a > b;
          case 7: return a <= b;
          case 8: return a >= b;
          case 9: return a == b;
          case 10: return a != b;
          case 11: return (a != 0) && (b != 0);
          case 12: return (a != 0) || (b != 0);
          case 13: return mu_round(a) && mu_round(b);  // reference :135-137
          case 14: return mu_round(a) || mu_round(b);  // reference :131-133
        }
        throw ExprError("internal: unknown operator " + n.op);
      }
      case NodeKind::Call: {
        double a = eval_node(*n.args[0], v);
        switch (n.code) {
          case 15: return std::sin(a);
          case 16: return std::cos(a);
          case 17: return std::tan(a);
          case 18: return std::asin(a);
          case 19: return std::acos(a);
          case 20: return std::atan(a);
          case 21: return std::sinh(a);
          case 22: return std::cosh(a);
          case 23: return std::tanh(a);
          case 24: return std::asinh(a);
          case 25: return std::acosh(a);
          case 26: return std::atanh(a);
          case 27: return std::log2(a);
          case 28: return std::log10(a);
          case 29: case 30: return std::log(a);
          case 31: return std::exp(a);
          case 32: return std::sqrt(a);
          case 33: return (a < 0) ? -1.0 : (a > 0) ? 1.0 : 0.0;
          case 34: return std::floor(a + 0.5);
          case 35: return std::fabs(a);
          case 36: return static_cast<double>(mu_round(a));
          case 37: return std::ceil(a);
          case 38: return std::floor(a);
          case 39: return 1.0 / std::tan(a);
          case 40: return 1.0 / std::sin(a);
          case 41: return 1.0 / std::cos(a);
          case 42: return std::erfc(a);
          case 43: return std::pow(a, eval_node(*n.args[1], v));
          case 44: return std::atan2(a, eval_node(*n.args[1], v));
          case 45: return mu_round(a) ? eval_node(*n.args[1], v) : eval_node(*n.args[2], v);
          case 46: case 47: case 48: case 49: {
            double r = a;
            for (size_t i = 1; i < n.args.size(); ++i) {
              double b = eval_node(*n.args[i], v);
              if (n.code == 46) r = std::min(r, b);
              else if (n.code == 47) r = std::max(r, b);
              else r += b;
            }
            if (n.code == 49) r /= static_cast<double>(n.args.size());
            return r;
          }
        }
        throw ExprError("internal: unknown function " + n.op);
      }
    }
    return 0;
  }

  // muParser lowers x^2, x^3, x^4 to repeated multiplication and calls
  // std::pow otherwise; both agree to the last bit or two, far inside 1e-10.
  static double pow_like_muparser(double a, double b) {
    if (b == 2.0) return a * a;
    if (b == 3.0) return a * a * a;
    if (b == 4.0) return a * a * a * a;
    return std::pow(a, b);
  }
};

// Split "a;b;c;" into components, dropping blanks-only pieces at the end
// (deal.II Utilities::split_string_list semantics for a trailing separator).
inline std::vector<std::string> split_components(const std::string& s) {
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) {
    if (c == ';') {
      out.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  bool blank = true;
  for (char c : cur)
    if (c != ' ' && c != '\t') blank = false;
  if (!blank) out.push_back(cur);
  return out;
}

// A vector-valued function of (x0,x1,y0,y1[,t]) with an optional mapper, the
// oracle twin of MultiscaleFunctionParser<2>.
class MultiscaleFunction {
 public:
  int n_components = 1;
  bool time_dependent = false;
  double time = 0.0;
  const MultiscaleFunction* mapper = nullptr;
  std::vector<Expression> comps;
  double macro_point[2] = {0, 0};

  // reference: multiscale_function_parser.cpp:54-107, :327-335
  void initialize(const std::string& variables, const std::string& expression,
                  const std::map<std::string, double>& constants, const MultiscaleFunction* mapper_ = nullptr,
                  bool time_dependent_ = false, int n_components_ = 1) {
    n_components = n_components_;
    mapper = mapper_;
    time_dependent = time_dependent_;
    std::vector<std::string> vars;
    std::string cur;
    for (char c : variables) {
      if (c == ',') {
        if (!cur.empty()) vars.push_back(cur);
        cur.clear();
      } else if (c != ' ') {
        cur.push_back(c);
      }
    }
    if (!cur.empty()) vars.push_back(cur);
    if (vars.size() != static_cast<size_t>(time_dependent ? 5 : 4)) throw ExprError("Wrong number of variables");
    auto parts = split_components(expression);
    if (static_cast<int>(parts.size()) != n_components)
      throw ExprError("expression has " + std::to_string(parts.size()) + " components, expected " +
                      std::to_string(n_components) + ": <" + expression + ">");
    comps.clear();
    for (auto& p : parts) comps.emplace_back(p, vars, constants);
  }
  void set_time(double t) { time = t; }
  void set_macro_point(const double* x) {
    macro_point[0] = x[0];
    macro_point[1] = x[1];
  }
  // reference :338-375
  double mvalue(const double* px, const double* py, int component = 0) const {
    double v[5];
    v[0] = px[0];
    v[1] = px[1];
    if (mapper) {
      double m[2];
      mapper->mmap(px, py, m);
      v[2] = m[0];
      v[3] = m[1];
    } else {
      v[2] = py[0];
      v[3] = py[1];
    }
    v[4] = time;
    return comps[component].eval(v);
  }
  // reference :377-381
  double value(const double* py, int component = 0) const { return mvalue(macro_point, py, component); }
  // reference :383-405 (row-major dim x dim)
  void mtensor_value(const double* px, const double* py, double F[4]) const {
    double v[5] = {px[0], px[1], py[0], py[1], time};
    for (int i = 0; i < 4; ++i) F[i] = comps[i].eval(v);
  }
  // reference :407-430
  void mmap(const double* px, const double* py, double out[2]) const {
    double v[5] = {px[0], px[1], py[0], py[1], time};
    out[0] = comps[0].eval(v);
    out[1] = comps[1].eval(v);
  }
};

// Macro functions of (x0,x1[,t]) — deal.II FunctionParser<2> as used in pde_data.cpp.
class MacroFunction {
 public:
  Expression e;
  bool time_dependent = false;
  double time = 0;
  void initialize(const std::string& variables, const std::string& expression,
                  const std::map<std::string, double>& constants, bool time_dependent_ = false) {
    time_dependent = time_dependent_;
    std::vector<std::string> vars;
    std::string cur;
    for (char c : variables) {
      if (c == ',') {
        if (!cur.empty()) vars.push_back(cur);
        cur.clear();
      } else if (c != ' ') {
        cur.push_back(c);
      }
    }
    if (!cur.empty()) vars.push_back(cur);
    auto parts = split_components(expression);
    if (parts.size() != 1) throw ExprError("macro function must have one component");
    e.parse(parts[0], vars, constants);
  }
  void set_time(double t) { time = t; }
  double value(const double* p) const {
    double v[3] = {p[0], p[1], time};
    return e.eval(v);
  }
};

}  // namespace orc
