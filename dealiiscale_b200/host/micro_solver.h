// MicroSolver: the reference's micro-solver interface, backed by libmsgpu through its C ABI.
//
// Same public surface as the reference classes (src/elliptic-elliptic/micro.h:54-98,
// src/bio-multiscale/micro.h:66-104, src/elliptic-parabolic/rho_solver.h:53-100):
// set_grid_locations, setup, set_macro_solution, run / assemble_and_solve_all / iterate,
// compute_all_errors, compute_all_residuals, get_num_grids, solutions.  The reference borrows
// pointers to the macro vectors and reads them at every run(); so does this class — the
// values are shipped to the device when run() starts.  There is no CPU fallback: if the CUDA
// library cannot create a context the constructor throws.
#pragma once
#include <array>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/msgpu.h"
#include "distributed.h"
#include "macro_fem.h"
#include "prm.h"

namespace ms {

class MicroSolver {
 public:
  MicroSolver(int problem, int cells_per_axis, int q_degree, int mesh_kind, int device = -1)
      : problem_(problem), cells_(cells_per_axis) {
    msgpu_desc d{problem, mesh_kind, cells_per_axis, q_degree, device};
    const int rc = msgpu_create(&ctx_, &d);
    if (rc != MSGPU_OK) throw std::runtime_error("msgpu_create failed: no usable CUDA device (there is no CPU fallback)");
    n_dofs_ = (cells_per_axis + 1) * (cells_per_axis + 1);
  }
  ~MicroSolver() { msgpu_destroy(ctx_); }
  MicroSolver(const MicroSolver&) = delete;
  MicroSolver& operator=(const MicroSolver&) = delete;

  void set_function(int slot, const std::string& expr, const Constants& c, bool time_dependent = false) {
    std::vector<const char*> names;
    std::vector<double> values;
    for (const auto& kv : c) {
      names.push_back(kv.first.c_str());
      values.push_back(kv.second);
    }
    check(msgpu_set_function(ctx_, slot, expr.c_str(), names.data(), values.data(), static_cast<int>(names.size()),
                             time_dependent));
  }
  void set_scalars(const std::vector<double>& v) { check(msgpu_set_scalars(ctx_, v.data(), static_cast<int>(v.size()))); }

  // ---- the reference interface ----
  // All N macro support points.  With several ranks (distributed.h) this process keeps the contiguous block
  // [k0, k1) of msgpu_partition; every per-grid array the MACRO side sees (coupling integrals, error norms)
  // still has N entries, allgathered over NCCL / NVLink.
  void set_grid_locations(const std::vector<std::array<double, 2>>& locations) {
    n_total_ = static_cast<int>(locations.size());
    Distributed& d = distributed();
    if (d.sits_out()) throw std::runtime_error("this rank sits the level out (Distributed::plan_level)");
    ranks_ = d.active() ? d.ranks_in_use() : 1;
    k0_ = 0;
    k1_ = n_total_;
    if (ranks_ > 1) check(msgpu_partition(n_total_, ranks_, d.rank, &k0_, &k1_));
    num_grids_ = k1_ - k0_;
    if (num_grids_ < 1) throw std::runtime_error("empty block of macro points: call Distributed::plan_level first");
    check(msgpu_set_grid_locations(ctx_, &locations[k0_][0], num_grids_));
  }
  // Whether the fused coupling + allgather over peer memory is used (up to 8 ranks; MSGPU_NCCL_ALLGATHER=1: NCCL).
  static bool wants_p2p(int ranks) {
    const char* nccl_only = std::getenv("MSGPU_NCCL_ALLGATHER");
    return ranks > 1 && ranks <= 8 && !(nccl_only && nccl_only[0] == '1');
  }
  void setup() {
    Distributed& d = distributed();
    if (d.active()) {  // ncclGetUniqueId on rank 0 -> everybody -> ncclCommInitRank over the ranks in use
      char id[128] = {};
      if (d.rank == 0 && ranks_ > 1 && !d.share_device) check(msgpu_comm_unique_id(id));
      d.broadcast(id, sizeof id, 0);
      // ranks sharing one device: no NCCL communicator (id = NULL), the peer-memory exchange must follow
      if (ranks_ > 1) check(msgpu_comm_init(ctx_, d.share_device ? nullptr : id, d.rank, ranks_, n_total_, k0_));
    }
    check(msgpu_setup(ctx_));
    if (d.active() && wants_p2p(d.ranks_in_use())) {
      // fused coupling + allgather over peer memory (csrc/comm.cu): exchange the 64-byte IPC handles.  If CUDA IPC is
      // refused anywhere (export or attach), every rank drops back to the NCCL allgather together.
      unsigned char mine[65] = {};  // [0]: this rank could export, [1..64]: its handle
      const char* fail = std::getenv("MSGPU_TEST_FAIL_P2P");  // test hook: that rank pretends CUDA IPC is refused
      const bool pretend = fail && std::atoi(fail) == d.rank;
      if (ranks_ > 1 && !pretend) mine[0] = msgpu_comm_p2p_export(ctx_, mine + 1) == MSGPU_OK ? 1 : 0;
      std::vector<unsigned char> all;
      d.allgather(mine, sizeof mine, all);
      bool everybody = true;
      std::vector<unsigned char> handles(static_cast<size_t>(ranks_) * 64);
      for (int r = 0; r < ranks_; ++r) {
        everybody = everybody && all[static_cast<size_t>(r) * 65] == 1;
        std::copy(all.begin() + static_cast<size_t>(r) * 65 + 1, all.begin() + static_cast<size_t>(r + 1) * 65,
                  handles.begin() + static_cast<size_t>(r) * 64);
      }
      unsigned char attached = 0;
      if (ranks_ > 1 && everybody) attached = msgpu_comm_p2p_attach(ctx_, handles.data()) == MSGPU_OK ? 1 : 0;
      d.allgather(&attached, 1, all);
      bool all_attached = everybody;
      for (int r = 0; r < ranks_; ++r) all_attached = all_attached && all[r] == 1;
      if (ranks_ > 1 && !all_attached) {
        if (d.share_device) throw std::runtime_error("ranks share a device and CUDA IPC is unavailable: no exchange path");
        check(msgpu_comm_p2p_detach(ctx_));
        if (d.rank == 0) std::fprintf(stderr, "msgpu: peer-memory exchange unavailable, using ncclAllGather\n");
      }
    }
  }
  // A rank that sits a level out performs the same out-of-band exchanges as the ranks that work on it.
  static void idle_setup() {
    Distributed& d = distributed();
    if (!d.active()) return;
    char id[128] = {};
    d.broadcast(id, sizeof id, 0);
    if (wants_p2p(d.ranks_in_use())) {
      unsigned char mine[65] = {}, attached = 0;
      std::vector<unsigned char> all;
      d.allgather(mine, sizeof mine, all);
      d.allgather(&attached, 1, all);
    }
  }
  int first_grid() const { return k0_; }           // global index of this rank's first micro system
  int n_total() const { return n_total_; }
  bool owns(int k) const { return k >= k0_ && k < k1_; }
  void set_macro_solution(const std::vector<double>* u, const std::vector<double>* w = nullptr) {
    macro_u_ = u;
    macro_w_ = w;
  }
  unsigned int get_num_grids() const { return n_total_; }  // all ranks' grids, as the macro solver counts them
  unsigned int get_num_local_grids() const { return num_grids_; }
  int n_dofs() const { return n_dofs_; }
  int cells_per_axis() const { return cells_; }

  void assemble_system() {
    push_macro();
    check(msgpu_assemble(ctx_));
  }
  // elliptic, bio: starts the macro-independent part of the next assemble_system() on a second stream, so that it runs beside
  // the macro solve (msgpu_assemble_begin; a no-op for the other problems and with MSGPU_OVERLAP_ASSEMBLY=0)
  void assemble_begin() { check(msgpu_assemble_begin(ctx_)); }
  void solve() {
    iterations.assign(num_grids_, 0);
    check(msgpu_solve(ctx_, 1e-12, problem_ == MSGPU_ELLIPTIC ? 1000 : 10000, precond, iterations.data(), nullptr));
  }
  void run() {  // micro.cpp:218-221
    assemble_system();
    solve();
  }
  void assemble_and_solve_all() { run(); }  // bio micro.cpp:292-305
  void iterate(double dt, double t) {       // rho_solver.cpp:232-237
    push_macro();
    iterations.assign(num_grids_, 0);
    check(msgpu_time_step(ctx_, dt, t, 1e-12, 10000, iterations.data()));
  }
  // MacroSolver::compute_microscopic_contribution reads these
  void coupling(std::vector<double>& c0, std::vector<double>* c1 = nullptr) {
    c0.assign(n_total_, 0.0);
    std::vector<double> dummy(n_total_);
    if (c1) c1->assign(n_total_, 0.0);
    check(msgpu_coupling(ctx_, c0.data(), c1 ? c1->data() : dummy.data()));
  }
  // per-grid scalars of all ranks: local[f][num_grids] -> all[f][n_total].  NCCL (msgpu_comm_allgather); ranks that
  // share a device have no communicator and go through the host exchange
  void gather_scalars(const double* local, double* all, int fields) {
    Distributed& d = distributed();
    if (!(d.active() && d.share_device && ranks_ > 1)) {
      check(msgpu_comm_allgather(ctx_, local, all, fields));
      return;
    }
    if (ranks_ != d.world) throw std::runtime_error("ranks that share a device must all take part in every level");
    const int B = (n_total_ + ranks_ - 1) / ranks_;  // block of msgpu_partition
    std::vector<double> mine(static_cast<size_t>(fields) * B, 0.0);
    for (int f = 0; f < fields; ++f)
      std::copy(local + static_cast<size_t>(f) * num_grids_, local + static_cast<size_t>(f + 1) * num_grids_,
                mine.begin() + static_cast<size_t>(f) * B);
    std::vector<unsigned char> blob;
    d.allgather(mine.data(), mine.size() * sizeof(double), blob);
    const double* got = reinterpret_cast<const double*>(blob.data());
    for (int r = 0; r < ranks_; ++r)
      for (int f = 0; f < fields; ++f)
        for (int i = 0; i < B && r * B + i < n_total_; ++i)
          all[static_cast<size_t>(f) * n_total_ + r * B + i] = got[(static_cast<size_t>(r) * fields + f) * B + i];
  }
  // per-grid norms; the aggregation over the macro mesh is done by the caller (micro.cpp:206-214)
  void grid_errors(std::vector<double>& l2, std::vector<double>& h1) {
    std::vector<double> local(2 * static_cast<size_t>(num_grids_)), all(2 * static_cast<size_t>(n_total_));
    check(msgpu_errors(ctx_, local.data(), local.data() + num_grids_));
    gather_scalars(local.data(), all.data(), 2);  // a copy on one rank
    l2.assign(all.begin(), all.begin() + n_total_);
    h1.assign(all.begin() + n_total_, all.end());
  }
  void grid_residuals(std::vector<double>& l2) {
    std::vector<double> local(num_grids_);
    l2.assign(n_total_, 0.0);
    check(msgpu_residuals(ctx_, local.data()));
    gather_scalars(local.data(), l2.data(), 1);
  }
  // ---- the stop test on the device (include/msgpu.h, msgpu_norm_*): the per-grid norms never reach the host.
  // Off with MSGPU_HOST_STOP_TEST=1 (A/B) and for ranks that share a device (no NCCL communicator).
  void norm_setup(const MacroMesh& macro_mesh) {
    const char* host = std::getenv("MSGPU_HOST_STOP_TEST");
    Distributed& d = distributed();
    if ((host && host[0] == '1') || (d.active() && d.share_device)) return;
    check(msgpu_norm_setup(ctx_, &macro_mesh.cell_dofs[0][0], static_cast<int>(macro_mesh.cell_dofs.size()), macro_mesh.n,
                           macro_mesh.h));
    norms_on_device_ = true;
  }
  bool norms_on_device() const { return norms_on_device_; }
  double norm_of_values(const std::vector<double>& v, int q) {
    double out = 0;
    check(msgpu_norm_of_values(ctx_, v.data(), q, &out));
    return out;
  }
  void compute_all_errors(const MacroMesh& macro_mesh, int q, double& l2_error, double& h1_error) {
    if (norms_on_device_) {
      check(msgpu_norm_of_errors(ctx_, q, &l2_error, &h1_error));
      return;
    }
    std::vector<double> l2, h1;
    grid_errors(l2, h1);
    l2_error = nodal_l2_norm(macro_mesh, l2, q);
    h1_error = nodal_l2_norm(macro_mesh, h1, q);
  }
  void compute_all_residuals(const MacroMesh& macro_mesh, int q, double& l2_residual) {
    if (norms_on_device_) {
      check(msgpu_norm_of_residuals(ctx_, q, &l2_residual));
      return;
    }
    std::vector<double> r;
    grid_residuals(r);
    l2_residual = nodal_l2_norm(macro_mesh, r, q);
  }
  // micro.cpp:224-234, bio micro.cpp:381-391, rho_solver.cpp:353-359 (t: parabolic only)
  void set_exact_solution(double t = 0.0) { check(msgpu_set_exact_solution(ctx_, t)); }
  // `solutions` of the reference (micro.h:97), reference DoF numbering
  // k: global grid index; the grid must live on this rank (owns(k))
  std::vector<double> solution(int k) {
    std::vector<double> out(n_dofs_);
    check(msgpu_get_solutions(ctx_, k - k0_, k - k0_ + 1, out.data()));
    return out;
  }
  void solutions_range(int k0, int k1, std::vector<double>& out) {
    out.assign(static_cast<size_t>(k1 - k0) * n_dofs_, 0.0);
    check(msgpu_get_solutions(ctx_, k0 - k0_, k1 - k0_, out.data()));
  }
  void all_solutions(std::vector<double>& out) {
    out.assign(static_cast<size_t>(num_grids_) * n_dofs_, 0.0);
    check(msgpu_get_solutions(ctx_, 0, num_grids_, out.data()));
  }
  // ---- the macro system(s) on the device (include/msgpu.h, msgpu_macro_*) ----
  void macro_setup(int n_fields, int mesh_kind, int cells_per_axis) {
    check(msgpu_macro_setup(ctx_, n_fields, mesh_kind, cells_per_axis));
    macro_fields_ = n_fields;
  }
  void macro_set_system(int field, int source, const MacroMesh& mesh, const CsrMatrix& A, const CsrMatrix* C,
                        const std::vector<double>& b0, const std::vector<double>& x0) {
    check(msgpu_macro_set_system(ctx_, field, source, mesh.rowstart.data(), mesh.col.data(), A.val.data(),
                                 C ? C->val.data() : nullptr, b0.data(), x0.data()));
  }
  void macro_set_rhs(int field, const std::vector<double>& b) { check(msgpu_macro_set_rhs(ctx_, field, b.data())); }
  void macro_set_product_load(int field, double theta, bool nonlinear, const std::vector<double>& old0,
                              const std::vector<unsigned char>& dirichlet) {
    check(msgpu_macro_set_product_load(ctx_, field, theta, nonlinear ? 1 : 0, old0.data(), dirichlet.data()));
  }
  void macro_set_solution(int field, const std::vector<double>& x) { check(msgpu_macro_set_solution(ctx_, field, x.data())); }
  void macro_get_solution(int field, std::vector<double>& x) { check(msgpu_macro_get_solution(ctx_, field, x.data())); }
  std::vector<int> macro_solve(double tol, int max_it, bool adopt) {
    std::vector<int> its(macro_fields_, 0);
    check(msgpu_macro_solve(ctx_, tol, max_it, adopt ? 1 : 0, its.data()));
    return its;
  }
  // The macro step on several GPUs.  Default: every rank solves the (tiny) macro system itself from the
  // allgathered integrals.  MSGPU_MACRO_BROADCAST=1: rank 0 alone solves and the solution is broadcast with
  // ncclBroadcast (BASELINE.json north_star; msgpu_broadcast_macro also places this rank's block on the device).
  static bool macro_broadcast() {
    const char* e = std::getenv("MSGPU_MACRO_BROADCAST");
    return distributed().active() && distributed().ranks_in_use() > 1 && e && e[0] == '1';
  }
  static bool solves_macro_here() { return !macro_broadcast() || distributed().rank == 0; }
  void share_macro_solution(std::vector<double>& u, std::vector<double>* w = nullptr) {
    if (macro_broadcast()) check(msgpu_broadcast_macro(ctx_, u.data(), w ? w->data() : nullptr, 0));
  }
  msgpu_ctx* context() { return ctx_; }
  std::vector<int> iterations;  // CG steps of every grid in the last solve
  int precond = MSGPU_PRECOND_IDENTITY;  // the reference's choice (micro.cpp:179); Jacobi / SSOR are opt-ins

 private:
  msgpu_ctx* ctx_ = nullptr;
  int problem_, cells_, n_dofs_ = 0, num_grids_ = 0, macro_fields_ = 0;
  int n_total_ = 0, k0_ = 0, k1_ = 0;  // all grids / this rank's block
  int ranks_ = 1;                      // ranks sharing this problem
  bool norms_on_device_ = false;
  const std::vector<double>*macro_u_ = nullptr, *macro_w_ = nullptr;
  void push_macro() {
    if (!macro_u_) throw std::runtime_error("set_macro_solution was not called");
    check(msgpu_set_macro_solution(ctx_, macro_u_->data() + k0_, macro_w_ ? macro_w_->data() + k0_ : nullptr));
  }
  void check(int rc) {
    if (rc != MSGPU_OK) throw std::runtime_error(std::string("msgpu: ") + msgpu_last_error(ctx_));
  }
};

// deal.II ConvergenceTable::write_text (table_with_headers): centred keys, right-aligned entries.
inline std::string format_table(const std::vector<std::string>& keys, const std::vector<std::vector<std::string>>& rows) {
  std::vector<size_t> width(keys.size());
  for (size_t c = 0; c < keys.size(); ++c) {
    width[c] = keys[c].size();
    for (const auto& r : rows) width[c] = std::max(width[c], r[c].size());
  }
  std::string out;
  for (size_t c = 0; c < keys.size(); ++c) {
    const size_t front = (width[c] - keys[c].size()) / 2, back = width[c] - keys[c].size() - front;
    out += std::string(front, ' ') + keys[c] + std::string(back, ' ') + " ";
  }
  out += "\n";
  for (const auto& r : rows) {
    for (size_t c = 0; c < keys.size(); ++c) out += std::string(width[c] - r[c].size(), ' ') + r[c] + " ";
    out += "\n";
  }
  return out;
}
inline std::string sci3(double v) {
  char buf[32];
  std::snprintf(buf, sizeof buf, "%.3e", v);
  return buf;
}

}  // namespace ms
