// solve_elliptic [input/<id>.prm] — same command line as the reference (manufactured.cpp:29-40):
// levels i = 2..5 with (macro refinement i-1, micro refinement i), table appended to
// results/<id>_convergence_table.txt.
#include <regex>

#include "elliptic.h"

int main(int argc, char* argv[]) {
  std::string input_path = "input/map_test.prm";
  if (argc == 2) {
    input_path = argv[1];
  } else if (argc > 2) {
    std::printf("Too many arguments\n");
    return 1;
  }
  std::smatch m;
  const std::string id = std::regex_search(input_path, m, std::regex("input/(.+).prm")) ? m[1].str() : "run";
  const std::string output_path = "results/" + id + "_convergence_table.txt";
  ms::Distributed& dist = ms::distributed();  // one process per GPU when RANK / WORLD_SIZE / LOCAL_RANK are set
  if (dist.rank != 0 && !std::freopen("/dev/null", "w", stdout)) return 3;
  ms::ensure_directory("results");  // the reference expects it to exist (populate.sh links it)
  if (dist.rank == 0) { std::ofstream truncate(output_path, std::ios::trunc); }
  try {
    for (unsigned int i = 2; i < 6; i++) {
      const int macro_nodes = (1 << (i - 1)) + 1;
      dist.plan_level(macro_nodes * macro_nodes);  // coarse levels have fewer macro points than a big job has ranks
      if (dist.sits_out()) {
        ms::MicroSolver::idle_setup();
        continue;
      }
      ms::elliptic::Manager manager(i - 1, i, input_path, output_path, dist.active() ? dist.local_rank : -1);
      if (const char* e = std::getenv("MSGPU_CYCLE_LIMIT")) manager.cycle_limit = std::atoi(e);  // extension: stop after K cycles
      manager.setup();
      manager.run();
    }
    dist.finalize();
  } catch (const std::exception& e) {
    ms::distributed().abort_marker();  // the other ranks stop waiting for this one
    std::fprintf(stderr, "solve_elliptic: %s\n", e.what());
    return 2;
  }
  return 0;
}
