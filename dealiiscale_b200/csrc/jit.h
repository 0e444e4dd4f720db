// Run-time specialisation of the assembly kernels: the expression bytecode of one problem is
// lowered to straight-line CUDA C, spliced into the cell-moment kernel and compiled for sm_100a
// with NVRTC when the problem is set up.
//
// Replaces, like the VM (expr_vm.cuh), the muParser evaluation the reference performs at every
// quadrature point (src/tools/multiscale_function_parser.cpp:362, :401, :427) inside
// compute_pullback_objects / integrate_cell (src/elliptic-elliptic/micro.cpp:75-136) and
// local_assemble_system (src/bio-multiscale/micro.cpp:183-221).  Every bytecode instruction still
// becomes exactly one IEEE-754 operation (__dadd_rn, __dmul_rn, ... are never contracted), so the
// specialised kernel computes bit-for-bit what the interpreter computes; what disappears is the
// fetch / decode / dispatch work, which is >80 % of the interpreter's instructions.
//
// NVRTC is bound lazily with dlopen; when it is missing (or MSGPU_JIT=0) the interpreter kernels
// are used and msgpu_stats.jit_kernels stays 0.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "kernels.h"

namespace msgpu {

struct JitKernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  bool unrolled_rows = false;
};

// True when NVRTC could be loaded; otherwise *why says what was tried.
bool jit_available(std::string* why);

// CUDA C for one bytecode program evaluated for W lanes at once:
//   __device__ void <name>(const double* U, const double* L0, int l0s, const double* L1, int l1s,
//                          double (&R)[n_regs][W])
// LINE0 operand idx of lane e is L0[idx*l0s + e]; registers are R[idx][e].
std::string vm_to_cuda(const VmProgram& prog, const std::string& name, int W);

// Source of the cell-moment kernel specialised for these programs, this quadrature and this mode.
std::string cell_moments_source(const MomentPrograms& mp, int q, const double* qtab, const double* qsum,
                                bool* unrolled_rows);

// Source of the per-cell error kernel (exact solution program, MODE_POINTS) — mirrors body_cell_errors.
std::string cell_errors_source(const VmProgram& prog);

// Compiles `src` (entry point `kernel_name`, extern "C") for sm_100a and loads it on the current
// device.  Identical sources are compiled once per process.  Returns 0 on success; otherwise *log
// holds the NVRTC / loader message.
int jit_compile(const std::string& src, const char* kernel_name, JitKernel* out, std::string* log);

int launch_cell_moments_jit(cudaStream_t s, const JitKernel& jk, MeshDims md, Tables tb, int k0, int nk,
                            double stiffness_scale, double* cell, int* error_flag);

int launch_cell_errors_jit(cudaStream_t s, const JitKernel& jk, MeshDims md, const double* slots, int n_slots, int qe,
                           const double* pt, const double* wt, const double* x, int N, double fd, double* cellred);

}  // namespace msgpu
