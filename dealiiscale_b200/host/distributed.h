// One process per GPU for the C++ drivers: which rank this process is and a small
// out-of-band exchange for the two things the ranks must agree on before the first kernel — the 128-byte
// NCCL unique id and the 64-byte peer-memory handles of msgpu_comm_p2p_export.
//
// No reference call site: the reference is a single process (manager.cpp:25-26 hands raw pointers from the
// micro to the macro solver).  Launch as `torchrun --no-python --nproc-per-node N solve_biomath …` or any
// launcher that sets RANK, WORLD_SIZE and LOCAL_RANK; without them the drivers run on one GPU as before.
// The exchange goes through small files in a directory all ranks of the node can see (MSGPU_RENDEZVOUS_DIR,
// default /tmp/msgpu-rdzv-<parent pid>-<MASTER_PORT>): written to a temporary name and renamed, so a
// reader never sees a partial file.  Everything on the data path itself is NCCL / NVLink (csrc/comm.cu).
#pragma once
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace ms {

class Distributed {
 public:
  int rank = 0, world = 1, local_rank = 0;
  // MSGPU_RANKS_SHARE_DEVICE=1 (tests on a one-GPU box): every rank uses device 0.  NCCL refuses two ranks on one
  // device, so there is no communicator: the coupling integrals travel through peer memory (CUDA IPC works between
  // processes on one device) and the per-grid scalars of the stop test through the host exchange below.
  bool share_device = false;
  // Ranks that take part in the CURRENT problem (set by the driver before it builds a Manager).  The partition
  // gives rank r the block [r B, min((r+1) B, N)) with B = ceil(N / R); with few macro points and many ranks the
  // last blocks would be empty (N = 9 on 4 ranks: 3 + 3 + 3 + 0), so a coarse level runs on the largest R <= world
  // that leaves nobody empty and the other ranks sit the level out (MicroSolver::idle_setup keeps them in step
  // with the out-of-band exchanges).
  int level_ranks = 0;
  int ranks_in_use() const { return level_ranks > 0 ? level_ranks : world; }
  bool sits_out() const { return rank >= ranks_in_use(); }
  static int usable_ranks(int n_total, int world) {
    int R = world < 1 ? 1 : world;
    while (R > 1 && static_cast<long long>(R - 1) * ((n_total + R - 1) / R) >= n_total) --R;
    return R;
  }
  void plan_level(int n_total) { level_ranks = usable_ranks(n_total, world); }

  static Distributed from_env() {
    Distributed d;
    const char* r = std::getenv("RANK");
    const char* w = std::getenv("WORLD_SIZE");
    const char* l = std::getenv("LOCAL_RANK");
    if (r && w && std::atoi(w) > 1) {
      d.rank = std::atoi(r);
      d.world = std::atoi(w);
      d.local_rank = l ? std::atoi(l) : d.rank;
      const char* share = std::getenv("MSGPU_RANKS_SHARE_DEVICE");
      d.share_device = share && share[0] == '1';
      if (d.share_device) d.local_rank = 0;
      const char* dir = std::getenv("MSGPU_RENDEZVOUS_DIR");
      const char* port = std::getenv("MASTER_PORT");
      // default name: unique per launch (the launcher's pid and run id), private to the user
      const char* run_id = std::getenv("TORCHELASTIC_RUN_ID");
      d.dir_ = dir ? dir
                   : "/tmp/msgpu-rdzv-" + std::to_string(static_cast<long>(getuid())) + "-" +
                         std::to_string(static_cast<long>(getppid())) + "-" + (port ? port : "0") + "-" +
                         (run_id ? run_id : "run");
      std::error_code ec;
      std::filesystem::create_directories(d.dir_, ec);
      std::filesystem::permissions(d.dir_, std::filesystem::perms::owner_all, std::filesystem::perm_options::replace, ec);
    }
    return d;
  }
  bool active() const { return world > 1; }

  // allgather of one fixed-size blob per rank: all[r*bytes .. (r+1)*bytes) = blob of rank r
  void allgather(const void* mine, size_t bytes, std::vector<unsigned char>& all, double timeout_s = 300.0) {
    all.assign(bytes * world, 0);
    if (!active()) {
      std::copy(static_cast<const unsigned char*>(mine), static_cast<const unsigned char*>(mine) + bytes, all.begin());
      return;
    }
    const std::string tag = dir_ + "/x" + std::to_string(counter_++) + ".";
    {
      const std::string tmp = tag + std::to_string(rank) + ".tmp", fin = tag + std::to_string(rank);
      std::ofstream out(tmp, std::ios::binary);
      out.write(static_cast<const char*>(mine), static_cast<std::streamsize>(bytes));
      out.close();
      std::filesystem::rename(tmp, fin);
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < world; ++r) {
      const std::string f = tag + std::to_string(r);
      for (;;) {
        std::error_code ec;
        if (std::filesystem::exists(f, ec) && std::filesystem::file_size(f, ec) == bytes) break;
        if (std::filesystem::exists(dir_ + "/abort", ec))
          throw std::runtime_error("another rank failed (see its message); leaving the exchange");
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
          throw std::runtime_error("rendezvous timed out waiting for rank " + std::to_string(r) + " (" + f + ")");
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
      }
      std::ifstream in(f, std::ios::binary);
      in.read(reinterpret_cast<char*>(all.data() + bytes * r), static_cast<std::streamsize>(bytes));
    }
    // everybody has written exchange c, so everybody has finished reading exchange c-1: drop this rank's old file
    if (!last_file_.empty()) {
      std::error_code ec;
      std::filesystem::remove(last_file_, ec);
    }
    last_file_ = tag + std::to_string(rank);
  }
  // blob of rank `root` to everybody
  void broadcast(void* blob, size_t bytes, int root = 0) {
    std::vector<unsigned char> all;
    allgather(blob, bytes, all);
    std::copy(all.begin() + bytes * root, all.begin() + bytes * (root + 1), static_cast<unsigned char*>(blob));
  }
  // A rank that is about to die tells the others, so that they do not wait for the time-out.
  void abort_marker() {
    if (!active()) return;
    std::ofstream(dir_ + "/abort") << rank;
  }
  void barrier() {
    unsigned char b = 1;
    std::vector<unsigned char> all;
    allgather(&b, 1, all);
  }
  // Ends the exchange: after a last barrier every other rank leaves a marker and goes; rank 0 waits for the
  // markers (nobody reads anything any more) and removes the directory.
  void finalize() {
    if (!active()) return;
    barrier();
    if (rank != 0) {
      std::ofstream(dir_ + "/bye." + std::to_string(rank)) << 1;
      return;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 1; r < world; ++r) {
      std::error_code ec;
      while (!std::filesystem::exists(dir_ + "/bye." + std::to_string(r), ec) &&
             std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 30.0)
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
    std::error_code ec;
    std::filesystem::remove_all(dir_, ec);
  }

 private:
  std::string dir_, last_file_;
  unsigned long counter_ = 0;
};

// the one per process: every Manager of a sweep shares it, so exchange names never repeat
inline Distributed& distributed() {
  static Distributed d = Distributed::from_env();
  return d;
}

}  // namespace ms
