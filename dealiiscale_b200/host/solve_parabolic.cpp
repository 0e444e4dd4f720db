// solve_parabolic [input/<id>.prm] — same command line as the reference (two_scale.cpp:16-33, :77-92):
// seven space-time refinement levels h_inv = round(8*2^(i/2)), t_inv = 4*2^i.
#include <cmath>
#include <regex>

#include "parabolic.h"

int main(int argc, char* argv[]) {
  std::string input_path = "input/full_nonlinear.prm";
  int levels = 7;
  if (argc >= 2) input_path = argv[1];
  if (argc >= 3) levels = std::stoi(argv[2]);  // extension: fewer sweep levels
  std::smatch m;
  const std::string id = std::regex_search(input_path, m, std::regex("input/(.+).prm")) ? m[1].str() : "run";
  const std::string output_path = "results/" + id + "_convergence_table.txt";
  ms::Distributed& dist = ms::distributed();  // one process per GPU when RANK / WORLD_SIZE / LOCAL_RANK are set
  if (dist.rank != 0 && !std::freopen("/dev/null", "w", stdout)) return 3;
  ms::ensure_directory("results");  // the reference expects it to exist (populate.sh links it)
  if (dist.rank == 0) { std::ofstream truncate(output_path, std::ios::trunc); }
  try {
    for (int i = 0; i < levels; i++) {
      const auto h_inv = static_cast<unsigned int>(std::round(8 * std::pow(2, i / 2.)));
      const auto t_inv = static_cast<unsigned int>(std::round(4 * std::pow(2, i)));
      dist.plan_level(static_cast<int>((h_inv + 1) * (h_inv + 1)));
      if (dist.sits_out()) {
        ms::MicroSolver::idle_setup();
        continue;
      }
      ms::parabolic::TimeManager manager(h_inv, h_inv, t_inv, input_path, output_path,
                                         dist.active() ? dist.local_rank : -1);
      manager.setup();
      manager.run();
    }
    dist.finalize();
  } catch (const std::exception& e) {
    ms::distributed().abort_marker();  // the other ranks stop waiting for this one
    std::fprintf(stderr, "solve_parabolic: %s\n", e.what());
    return 2;
  }
  return 0;
}
