// NVRTC binding, module loading and launch of the run-time specialised kernels (see jit.h).
#include "jit.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
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
    if (!ok) {
      n.tried = "libnvrtc lacks a required entry point";
      dlclose(n.handle);
      n.handle = nullptr;
    }
  });
  return n;
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
  }
  JitKernel jk;
  cudaError_t e = cudaLibraryLoadData(&jk.lib, ce.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&jk.kernel, jk.lib, kernel_name);
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
