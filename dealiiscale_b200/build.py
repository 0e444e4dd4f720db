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

CUDA_SOURCES = ["msgpu.cu", "kernels_assemble.cu", "kernels_cg.cu", "kernels_cg_stencil.cu", "kernels_extra.cu", "kernels_norms.cu", "comm.cu", "macro.cu"]
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


def _compile_one(args):
    src, obj, verbose = args
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build_msgpu(force: bool = False, verbose: bool = False) -> str:
    """Compile libmsgpu.so for sm_100a (cross-compiles without a GPU).  One object per source, compiled in
    parallel and only when the source (or any header) is newer than its object."""
    from concurrent.futures import ThreadPoolExecutor

    files, deps = _sources()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(files + deps):
        return LIB
    objdir = os.path.join(BUILD, "obj")
    os.makedirs(objdir, exist_ok=True)
    newest_dep = _newest(deps)
    jobs, objs = [], []
    for src in files:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_dep):
            jobs.append((src, obj, verbose))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        for src, rc, out in pool.map(_compile_one, jobs):
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{out}")
            if verbose:
                print(out)
    res = subprocess.run([_nvcc(), "-shared", "-o", LIB] + objs + ["-ldl"], cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


CHECKED_LIB = os.path.join(BUILD, "libmsgpu_checked.so")
CHECKED_SOURCES = ("kernels_cg.cu", "kernels_cg_stencil.cu")


def build_checked(force: bool = False) -> str:
    """TEST INFRASTRUCTURE: libmsgpu_checked.so, the product's objects with the CG kernels compiled -DMSGPU_CHECKED
    (csrc/checked.cuh: every shared-, distributed-shared- and tensor-memory access of the resident CG kernels is
    compared with the window it may touch and traps otherwise).  tests/test_gpu_checked_build.py runs the kernels
    through it; nothing in the product loads it."""
    from concurrent.futures import ThreadPoolExecutor

    build_msgpu()
    files, deps = _sources()
    objdir = os.path.join(BUILD, "obj")
    newest_dep = _newest(deps)
    objs, jobs = [], []
    for src in files:
        base = os.path.basename(src)
        obj = os.path.join(objdir, base + ".o")
        if base in CHECKED_SOURCES:
            obj = os.path.join(objdir, base + ".checked.o")
            if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_dep):
                jobs.append([_nvcc()] + NVCC_FLAGS + ["-DMSGPU_CHECKED", "-c", src, "-o", obj])
        objs.append(obj)
    if not jobs and os.path.exists(CHECKED_LIB) and os.path.getmtime(CHECKED_LIB) >= os.path.getmtime(LIB):
        return CHECKED_LIB
    with ThreadPoolExecutor(max_workers=2) as pool:
        for res in pool.map(lambda c: subprocess.run(c, cwd=CSRC, capture_output=True, text=True), jobs):
            if res.returncode != 0:
                raise RuntimeError("nvcc failed (checked build):\n" + res.stdout + res.stderr)
    res = subprocess.run([_nvcc(), "-shared", "-o", CHECKED_LIB] + objs + ["-ldl"], cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed (checked build):\n" + res.stdout + res.stderr)
    return CHECKED_LIB


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
    build_checked(force=force)
    build_oracle()
    build_emul()
    build_host_drivers()


if __name__ == "__main__":
    import sys

    build_msgpu(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_checked(force="--force" in sys.argv)
    build_oracle()
    build_emul()
    build_host_drivers()
    print("built", LIB)
