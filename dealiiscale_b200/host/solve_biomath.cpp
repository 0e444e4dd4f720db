// solve_biomath [input.prm [macro_refinement [micro_refinement [num_threads]]]] — same command line
// and timing line as the reference (manufactured.cpp:39-90).  num_threads is accepted for
// compatibility; the micro path runs on the GPU.  Under a launcher that sets RANK / WORLD_SIZE / LOCAL_RANK
// (e.g. `torchrun --no-python --nproc-per-node 8 solve_biomath input/biocase.prm 6 6`) the macro points are
// sharded over one process per GPU (distributed.h); rank 0 prints and writes the table.
#include <chrono>
#include <ctime>
#include <regex>

#include "bio.h"

int main(int argc, char* argv[]) {
  std::string input_path = "input/linear.prm";
  int macro_refinement = 2, micro_refinement = 4, num_threads = 0;
  if (argc >= 2) input_path = argv[1];
  if (argc >= 3) macro_refinement = std::stoi(argv[2]);
  if (argc >= 4) micro_refinement = std::stoi(argv[3]);
  if (argc >= 5) num_threads = std::stoi(argv[4]);
  ms::Distributed& dist = ms::distributed();
  if (dist.rank != 0 && !std::freopen("/dev/null", "w", stdout)) return 3;
  const auto wall0 = std::chrono::steady_clock::now();
  const std::clock_t cpu0 = std::clock();
  const std::string output_path = "results/out.txt";
  ms::ensure_directory("results");  // the reference expects it to exist (populate.sh links it)
  if (dist.rank == 0) { std::ofstream truncate(output_path, std::ios::trunc); }
  try {
    const int macro_nodes = (1 << macro_refinement) + 1;
    dist.plan_level(macro_nodes * macro_nodes);  // a coarse macro mesh may have fewer points than the job has ranks
    if (dist.sits_out()) {
      ms::MicroSolver::idle_setup();
      dist.finalize();
      return 0;
    }
    ms::bio::Manager manager(1u << macro_refinement, 1u << micro_refinement, input_path, output_path, num_threads,
                             dist.active() ? dist.local_rank : -1);
    if (const char* e = std::getenv("MSGPU_CYCLE_LIMIT")) manager.cycle_limit = std::atoi(e);  // extension: stop after K cycles
    const auto t1 = std::chrono::steady_clock::now();
    manager.setup();
    const auto t2 = std::chrono::steady_clock::now();
    manager.run();
    const auto t3 = std::chrono::steady_clock::now();
    dist.finalize();
    if (std::getenv("MSGPU_DRIVER_TIMING"))
      std::printf("[timing] construct %.2f s, setup %.2f s, run + output %.2f s\n",
                  std::chrono::duration<double>(t1 - wall0).count(), std::chrono::duration<double>(t2 - t1).count(),
                  std::chrono::duration<double>(t3 - t2).count());
  } catch (const std::exception& e) {
    ms::distributed().abort_marker();  // the other ranks stop waiting for this one
    std::fprintf(stderr, "solve_biomath: %s\n", e.what());
    return 2;
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
  const double cpu = double(std::clock() - cpu0) / CLOCKS_PER_SEC;
  std::printf("Results: 'macro_refinement', 'micro_refinement', 'num_threads', 'wall_time', 'cpu_time'\n");
  std::printf("%d, %d, %d, %.3f, %.3f\n", macro_refinement, micro_refinement, num_threads, wall, cpu);
  if (num_threads != 0 && dist.rank == 0) {
    std::smatch m;
    const std::string id = std::regex_search(input_path, m, std::regex("input/(.+).prm")) ? m[1].str() : "run";
    std::ofstream t("results/" + id + "_timings.txt", std::ios::app);
    t << num_threads << "\t" << wall << "\t" << cpu << std::endl;
  }
  return 0;
}
