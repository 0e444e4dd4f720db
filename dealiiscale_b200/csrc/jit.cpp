// NVRTC binding, module loading and launch of the run-time specialised kernels (see jit.h).
#include "jit.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace msgpu {

namespace {

// ------------------------------------------------------------------------------ NVRTC, lazily
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* handle = nullptr;
  std::string tried;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  int version = 0;  // major * 1000 + minor
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
      n.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.handle) break;
      n.tried += std::string(n.tried.empty() ? "" : ", ") + nm;
    }
    if (!n.handle) return;
    bool ok = true;
    auto bind = [&](auto& fp, const char* sym) {
      fp = reinterpret_cast<std::remove_reference_t<decltype(fp)>>(dlsym(n.handle, sym));
      if (!fp) ok = false;
    };
    bind(n.CreateProgram, "nvrtcCreateProgram");
    bind(n.CompileProgram, "nvrtcCompileProgram");
    bind(n.DestroyProgram, "nvrtcDestroyProgram");
    bind(n.GetCUBINSize, "nvrtcGetCUBINSize");
    bind(n.GetCUBIN, "nvrtcGetCUBIN");
    bind(n.GetProgramLogSize, "nvrtcGetProgramLogSize");
    bind(n.GetProgramLog, "nvrtcGetProgramLog");
    bind(n.GetErrorString, "nvrtcGetErrorString");
    bind(n.Version, "nvrtcVersion");
    int major = 0, minor = 0;
    if (ok && n.Version(&major, &minor) == 0) n.version = major * 1000 + minor;
    if (!ok) {
      n.tried = "libnvrtc lacks a required entry point";
      dlclose(n.handle);
      n.handle = nullptr;
    }
  });
  return n;
}

// ------------------------------------------------------------------------------ cubins on disk
// NVRTC needs 0.4 - 2 s per specialised kernel; the cubin depends on nothing but the source text, the NVRTC version
// and the target, so it is kept in a per-user directory (MSGPU_JIT_CACHE, default $XDG_CACHE_HOME/msgpu or
// ~/.cache/msgpu, falling back to /tmp/msgpu-jit-<uid>; MSGPU_JIT_CACHE=off disables it) under the hash of exactly
// that.  Files are written to a temporary name and renamed, so a reader never sees a partial cubin; a file that does
// not load is ignored and overwritten.
std::string cache_dir() {
  const char* e = std::getenv("MSGPU_JIT_CACHE");
  if (e && (std::strcmp(e, "off") == 0 || std::strcmp(e, "0") == 0)) return "";
  std::string dir;
  if (e && e[0]) {
    dir = e;
  } else if (const char* x = std::getenv("XDG_CACHE_HOME")) {
    dir = std::string(x) + "/msgpu";
  } else if (const char* h = std::getenv("HOME")) {
    dir = std::string(h) + "/.cache/msgpu";
    mkdir((std::string(h) + "/.cache").c_str(), 0700);
  } else {
    dir = "/tmp/msgpu-jit-" + std::to_string(static_cast<long>(getuid()));
  }
  if (mkdir(dir.c_str(), 0700) != 0) {
    struct stat st {};
    if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode) || st.st_uid != getuid()) return "";
  }
  return dir;
}

std::string cache_key(const std::string& kernel_name, const std::string& src, int nvrtc_version) {
  // FNV-1a, 2 x 64 bit (two different offsets), over version, target, name and source
  unsigned long long h1 = 1469598103934665603ull, h2 = 1099511628211ull * 31 + 7;
  auto feed = [&](const std::string& t) {
    for (unsigned char c : t) {
      h1 = (h1 ^ c) * 1099511628211ull;
      h2 = (h2 ^ (c + 0x9e)) * 1099511628211ull;
    }
  };
  feed("nvrtc " + std::to_string(nvrtc_version) + " sm_100a c++17 lineinfo\n");
  feed(kernel_name + "\n");
  feed(src);
  char buf[40];
  std::snprintf(buf, sizeof buf, "%016llx%016llx", h1, h2);
  return buf;
}

bool cache_read(const std::string& path, std::vector<char>& cubin) {
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in) return false;
  const std::streamsize n = in.tellg();
  if (n < 64) return false;
  cubin.resize(static_cast<size_t>(n));
  in.seekg(0);
  in.read(cubin.data(), n);
  return static_cast<bool>(in) && std::memcmp(cubin.data(), "\x7f" "ELF", 4) == 0;
}

void cache_write(const std::string& path, const std::vector<char>& cubin) {
  const std::string tmp = path + "." + std::to_string(static_cast<long>(getpid())) + ".tmp";
  {
    std::ofstream out(tmp, std::ios::binary);
    if (!out) return;
    out.write(cubin.data(), static_cast<std::streamsize>(cubin.size()));
    if (!out) {
      unlink(tmp.c_str());
      return;
    }
  }
  if (rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
}

std::mutex g_cache_mutex;
struct CacheEntry {
  std::vector<char> cubin;
  std::unordered_map<int, JitKernel> per_device;
};
std::unordered_map<std::string, CacheEntry> g_cache;  // key: kernel name + source

}  // namespace

bool jit_available(std::string* why) {
  Nvrtc& n = nvrtc();
  if (!n.handle && why) *why = "NVRTC not found (tried " + n.tried + ")";
  return n.handle != nullptr;
}

int jit_compile(const std::string& src, const char* kernel_name, JitKernel* out, std::string* log) {
  Nvrtc& n = nvrtc();
  if (!n.handle) {
    if (log) *log = "NVRTC not found (tried " + n.tried + ")";
    return 1;
  }
  int device = 0;
  cudaGetDevice(&device);
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  CacheEntry& ce = g_cache[std::string(kernel_name) + "\n" + src];
  auto hit = ce.per_device.find(device);
  if (hit != ce.per_device.end()) {
    *out = hit->second;
    return 0;
  }
  std::string disk;
  bool from_disk = false;
  if (ce.cubin.empty()) {
    const std::string dir = cache_dir();
    if (!dir.empty()) {
      disk = dir + "/" + cache_key(kernel_name, src, n.version) + ".cubin";
      from_disk = cache_read(disk, ce.cubin);
      if (!from_disk) ce.cubin.clear();
    }
  }
  if (ce.cubin.empty()) {
    nvrtcProgram prog = nullptr;
    int rc = n.CreateProgram(&prog, src.c_str(), "msgpu_jit.cu", 0, nullptr, nullptr);
    if (rc != 0) {
      if (log) *log = std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc);
      return 1;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    rc = n.CompileProgram(prog, 3, opts);
    if (rc != 0) {
      size_t ls = 0;
      n.GetProgramLogSize(prog, &ls);
      std::string l(ls, '\0');
      if (ls) n.GetProgramLog(prog, &l[0]);
      if (log) *log = std::string("nvrtcCompileProgram: ") + n.GetErrorString(rc) + "\n" + l;
      n.DestroyProgram(&prog);
      return 1;
    }
    size_t sz = 0;
    n.GetCUBINSize(prog, &sz);
    ce.cubin.resize(sz);
    rc = n.GetCUBIN(prog, ce.cubin.data());
    n.DestroyProgram(&prog);
    if (rc != 0 || sz == 0) {
      ce.cubin.clear();
      if (log) *log = "nvrtcGetCUBIN failed";
      return 1;
    }
    if (!disk.empty()) cache_write(disk, ce.cubin);
  }
  JitKernel jk;
  cudaError_t e = cudaLibraryLoadData(&jk.lib, ce.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&jk.kernel, jk.lib, kernel_name);
  if (e != cudaSuccess && from_disk) {
    // a stale or damaged file: forget it and compile (one recursion: the entry is empty again and the file is gone)
    cudaGetLastError();
    ce.cubin.clear();
    unlink(disk.c_str());
    g_cache.erase(std::string(kernel_name) + "\n" + src);
    g_cache_mutex.unlock();
    const int rc = jit_compile(src, kernel_name, out, log);
    g_cache_mutex.lock();
    return rc;
  }
  if (e != cudaSuccess) {
    if (log) *log = std::string("loading the compiled kernel: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return 1;
  }
  ce.per_device[device] = jk;  // lives for the process: contexts share it
  *out = jk;
  return 0;
}

int launch_cell_moments_jit(cudaStream_t s, const JitKernel& jk, MeshDims md, Tables tb, int k0, int nk,
                            double stiffness_scale, double* cell, int* error_flag) {
  if (nk == 0) return 0;
  const long long work = static_cast<long long>(nk) * md.ncells;
  void* args[] = {&md, &tb, &k0, &nk, &stiffness_scale, &cell, &error_flag};
  const cudaError_t e = cudaLaunchKernel(reinterpret_cast<const void*>(jk.kernel),
                                         dim3(static_cast<unsigned>((work + 127) / 128)), dim3(128), args, 0, s);
  return e == cudaSuccess ? 1 : -1;
}

int launch_cell_errors_jit(cudaStream_t s, const JitKernel& jk, MeshDims md, const double* slots, int n_slots, int qe,
                           const double* pt, const double* wt, const double* x, int N, double fd, double* cellred) {
  if (N == 0) return 0;
  const long long work = static_cast<long long>(N) * md.ncells;
  void* args[] = {&md, &slots, &n_slots, &qe, &pt, &wt, &x, &N, &fd, &cellred};
  const cudaError_t e = cudaLaunchKernel(reinterpret_cast<const void*>(jk.kernel),
                                         dim3(static_cast<unsigned>((work + 127) / 128)), dim3(128), args, 0, s);
  return e == cudaSuccess ? 1 : -1;
}

}  // namespace msgpu
