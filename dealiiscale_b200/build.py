"""Build recipes: libmsgpu.so (the product, nvcc for sm_100a) and the two test-side libraries.

Everything is built IN-TREE (``dealiiscale_b200/_build``, ``oracle/_build``, ``tests/native/_build``)
so that the shared objects travel with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(BUILD, "libmsgpu.so")

CUDA_SOURCES = ["msgpu.cu", "kernels_assemble.cu", "kernels_cg.cu", "kernels_extra.cu", "comm.cu", "macro.cu"]
HOST_SOURCES = ["expr_compile.cpp", "problem_programs.cpp", "jit.cpp", "jit_codegen.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",  # host side only; device contraction is nvcc's default (-fmad=true)
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _newest(paths) -> float:
    return max(os.path.getmtime(p) for p in paths)


def _sources():
    files = [os.path.join(CSRC, f) for f in CUDA_SOURCES + HOST_SOURCES]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "msgpu.h"))
    return files, deps


def build_msgpu(force: bool = False, verbose: bool = False) -> str:
    """Compile libmsgpu.so for sm_100a (cross-compiles without a GPU)."""
    files, deps = _sources()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(files + deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + files + ["-ldl"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_oracle() -> str:
    """TEST INFRASTRUCTURE: the CPU restatement of the reference (oracle/)."""
    subprocess.check_call(["make", "-s"], cwd=os.path.join(ROOT, "oracle"))
    return os.path.join(ROOT, "oracle", "_build", "liboracle.so")


def build_emul() -> str:
    """TEST INFRASTRUCTURE: CPU harness around the kernel bodies (tests/native/)."""
    subprocess.check_call(["make", "-s"], cwd=os.path.join(ROOT, "tests", "native"))
    return os.path.join(ROOT, "tests", "native", "_build", "libemul.so")


def build_host_drivers() -> None:
    """The C++ host drivers (solve_elliptic, solve_biomath, solve_parabolic) on top of libmsgpu."""
    host = os.path.join(PKG, "host")
    if os.path.exists(os.path.join(host, "Makefile")):
        subprocess.check_call(["make", "-s"], cwd=host)


def build_all(force: bool = False) -> None:
    build_msgpu(force=force)
    build_oracle()
    build_emul()
    build_host_drivers()


if __name__ == "__main__":
    import sys

    build_msgpu(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle()
    build_emul()
    build_host_drivers()
    print("built", LIB)
