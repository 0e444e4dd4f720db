"""CPU tier: the code generator behind the run-time specialised cell-moment kernel (csrc/jit_codegen.cpp).

1. The generated CUDA C is compiled as HOST C++ behind a small shim (g++, no FMA contraction) and run
   cell by cell; its output must equal the interpreter body's (tests/native harness) bit for bit.
2. NVRTC (which needs no GPU) must accept the same text for sm_100a.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import INPUT, Emul, configure_bio, configure_elliptic, dptr

SHIM = r"""
#include <cmath>
#include <cstring>
#define __device__
#define __global__
#define __forceinline__ inline
#define __constant__ static const
#define __launch_bounds__(x)
struct Idx3 { unsigned x, y, z; };
static Idx3 blockIdx, threadIdx;
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
"""

DRIVER = r"""
extern "C" void host_run(const int* d, double h, const double* slots, const double* T0, const double* T1, int nk,
                         double scale, double* cell, int* flag) {
  MeshDims md{d[3], d[4], d[5], d[6], d[7], d[8], d[9], h};
  Tables tb{slots, T0, T1, d[0], d[1], d[2]};
  const long long work = static_cast<long long>(nk) * md.ncells;
  for (long long t = 0; t < work; ++t) {
    blockIdx.x = static_cast<unsigned>(t / 128);
    threadIdx.x = static_cast<unsigned>(t % 128);
    k_cell_moments_jit(md, tb, 0, nk, scale, cell, flag);
  }
}
"""


def _emul(problem, prm, nc, q):
    from dealiiscale_b200.prm import BioData, EllipticData

    if problem == "bio":
        e = Emul(1, 1, nc, q)
        configure_bio(e, BioData(os.path.join(INPUT, prm)))
    else:
        e = Emul(0, 0, nc, q)
        configure_elliptic(e, EllipticData(os.path.join(INPUT, prm)))
    xy = np.array([[-0.5, 0.25], [0.3, 0.7], [0.9, -0.9]])
    e.setup(xy)
    return e, xy


def _source(e):
    e.lib.emul_jit_source.restype = C.c_int
    e.lib.emul_jit_source.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    n = e.lib.emul_jit_source(e.h, None, 0)
    assert n > 0, "the volume program could not be lowered"
    buf = C.create_string_buffer(n)
    e.lib.emul_jit_source(e.h, buf, n)
    return buf.value.decode()


CASES = [
    ("bio", "biocase.prm", 3, 12, 1),             # const-F
    ("bio", "biocase-curved.prm", 3, 12, 1),      # const-F: the map bends with the macro position only
    ("elliptic", "elliptic_non_linear.prm", 2, 12, 0),
    ("elliptic", "rot_map_test.prm", 4, 12, 1),
    ("elliptic", "elliptic_linear.prm", 4, 3, 1),  # another quadrature
]


@pytest.mark.parametrize("problem,prm,nc,q,mode", CASES)
def test_generated_kernel_equals_interpreter_on_the_host(problem, prm, nc, q, mode, tmp_path):
    e, xy = _emul(problem, prm, nc, q)
    assert e.moment_mode() == mode
    src = _source(e)
    cpp = tmp_path / "k.cpp"
    cpp.write_text(SHIM + src + DRIVER)
    so = tmp_path / "k.so"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-Wno-unknown-pragmas", "-o", str(so), str(cpp)])
    lib = C.CDLL(str(so))

    # interpreter body (what k_cell_moments_vec runs per thread)
    N = len(xy)
    e.assemble(1.0 + xy[:, 0], 0.5 - xy[:, 1])
    ncells = nc * nc
    want = np.zeros(N * 18 * ncells)
    e.lib.emul_get_cell_scratch.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
    assert e.lib.emul_get_cell_scratch(e.h, dptr(want), want.size) == want.size

    slots, t0, t1 = C.c_void_p(), C.c_void_p(), C.c_void_p()
    dims = (C.c_int * 10)()
    e.lib.emul_table_pointers.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 3 + [C.POINTER(C.c_int)]
    e.lib.emul_table_pointers(e.h, C.byref(slots), C.byref(t0), C.byref(t1), dims)
    e.lib.emul_mesh_h.restype = C.c_double
    e.lib.emul_mesh_h.argtypes = [C.c_void_p]
    got = np.zeros_like(want)
    flag = C.c_int(0)
    scale = 1.0  # D_2 of the bio sets used here is 1, elliptic always 1
    lib.host_run.argtypes = [C.POINTER(C.c_int), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                             C.c_void_p, C.POINTER(C.c_int)]
    lib.host_run(dims, e.lib.emul_mesh_h(e.h), slots, t0, t1, N, scale, dptr(got), C.byref(flag))
    assert flag.value == 0
    assert np.abs(want).max() > 0
    got, want = got.reshape(N, 18, ncells), want.reshape(N, 18, ncells)
    if mode == 0:
        assert np.array_equal(got, want)
    else:
        # F constant per system: the generated kernel and the vectorised interpreter write the cell-independent
        # entries for cell 0 only (the gather reads them there, GatherArgs::uniform_stiffness); the scalar
        # interpreter, which integrates every cell numerically, shows that they really are the same everywhere
        uniform = [i for i in range(18) if i < 10 or i >= 14]
        assert np.array_equal(got[:, 10:14, :], want[:, 10:14, :])
        assert np.array_equal(got[:, uniform, 0], want[:, uniform, 0])
        e.set_vector_mode(False)
        e.assemble(1.0 + xy[:, 0], 0.5 - xy[:, 1])
        full = np.zeros(N * 18 * ncells)
        assert e.lib.emul_get_cell_scratch(e.h, dptr(full), full.size) == full.size
        full = full.reshape(N, 18, ncells)
        scale_ref = np.abs(full[:, uniform, :1]).max(axis=1, keepdims=True) + 1e-300
        assert (np.abs(full[:, uniform, :] - want[:, uniform, :1]) <= 1e-13 * scale_ref).all()
    e.close()


def _nvrtc():
    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    return None


@pytest.mark.parametrize("problem,prm", [("bio", "biocase.prm"), ("bio", "biocase-curved.prm"),
                                         ("elliptic", "elliptic_non_linear.prm")])
def test_nvrtc_accepts_the_generated_kernel(problem, prm):
    nv = _nvrtc()
    if nv is None:
        pytest.skip("NVRTC not installed")
    e, _ = _emul(problem, prm, 4, 12)
    src = _source(e)
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src.encode(), b"msgpu_jit.cu", 0, None, None) == 0
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo"]
    rc = nv.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts))
    size = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(size))
    log = C.create_string_buffer(size.value)
    nv.nvrtcGetProgramLog(prog, log)
    assert rc == 0, log.value.decode()
    nv.nvrtcGetCUBINSize(prog, C.byref(size))
    assert size.value > 1000
    nv.nvrtcDestroyProgram(C.byref(prog))
    e.close()


ERR_DRIVER = r"""
extern "C" void host_run_errors(const int* d, double h, const double* slots, int qe, const double* pt, const double* wt,
                                const double* x, int N, double fd, double* out) {
  MeshDims md{d[3], d[4], d[5], d[6], d[7], d[8], d[9], h};
  const long long work = static_cast<long long>(N) * md.ncells;
  for (long long t = 0; t < work; ++t) {
    blockIdx.x = static_cast<unsigned>(t / 128);
    threadIdx.x = static_cast<unsigned>(t % 128);
    k_cell_errors_jit(md, slots, d[0], qe, pt, wt, x, N, fd, out);
  }
}
"""


def _errors_source(e):
    e.lib.emul_jit_errors_source.restype = C.c_int
    e.lib.emul_jit_errors_source.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    n = e.lib.emul_jit_errors_source(e.h, None, 0)
    assert n > 0
    buf = C.create_string_buffer(n)
    e.lib.emul_jit_errors_source(e.h, buf, n)
    return buf.value.decode()


@pytest.mark.parametrize("problem,prm,nc", [("elliptic", "elliptic_non_linear.prm", 4), ("bio", "linear.prm", 3)])
def test_generated_error_kernel_equals_interpreter_on_the_host(problem, prm, nc, tmp_path):
    e, xy = _emul(problem, prm, nc, 12)
    src = _errors_source(e)
    cpp = tmp_path / "e.cpp"
    cpp.write_text(SHIM + src + ERR_DRIVER)
    so = tmp_path / "e.so"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-Wno-unknown-pragmas", "-o", str(so), str(cpp)])
    lib = C.CDLL(str(so))
    N, ncells = len(xy), nc * nc
    e.assemble(1.0 + xy[:, 0], 0.5 - xy[:, 1])
    e.solve(1e-12, 1000)
    want = np.zeros(N * 2 * ncells)
    e.lib.emul_cell_errors.argtypes = [C.c_void_p, C.c_void_p]
    assert e.lib.emul_cell_errors(e.h, dptr(want)) == 0
    slots, t0, t1 = C.c_void_p(), C.c_void_p(), C.c_void_p()
    dims = (C.c_int * 10)()
    e.lib.emul_table_pointers.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 3 + [C.POINTER(C.c_int)]
    e.lib.emul_table_pointers(e.h, C.byref(slots), C.byref(t0), C.byref(t1), dims)
    e.lib.emul_mesh_h.restype = C.c_double
    e.lib.emul_mesh_h.argtypes = [C.c_void_p]
    e.lib.emul_solution_pointer.restype = C.c_void_p
    e.lib.emul_solution_pointer.argtypes = [C.c_void_p]
    from helpers import load_oracle
    pt, wt = np.zeros(12), np.zeros(12)
    load_oracle().orc_gauss01(12, dptr(pt), dptr(wt))
    got = np.zeros_like(want)
    lib.host_run_errors.argtypes = [C.POINTER(C.c_int), C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    lib.host_run_errors(dims, e.lib.emul_mesh_h(e.h), slots, 12, dptr(pt), dptr(wt),
                        C.c_void_p(e.lib.emul_solution_pointer(e.h)), N, 1e-8, dptr(got))
    assert np.abs(want).max() > 0
    # the quadrature nodes come from two Gauss-Legendre routines that agree to the last ulp or two; the
    # central differences (step 1e-8) amplify that, so the H1 part is compared to 1e-6
    got, want = got.reshape(N, 2, ncells), want.reshape(N, 2, ncells)
    np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=1e-5, atol=1e-300)
    e.close()


def test_nvrtc_accepts_the_generated_error_kernel():
    nv = _nvrtc()
    if nv is None:
        pytest.skip("NVRTC not installed")
    e, _ = _emul("elliptic", "elliptic_non_linear.prm", 4, 12)
    src = _errors_source(e)
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src.encode(), b"msgpu_jit_err.cu", 0, None, None) == 0
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo"]
    rc = nv.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts))
    size = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(size))
    log = C.create_string_buffer(size.value)
    nv.nvrtcGetProgramLog(prog, log)
    assert rc == 0, log.value.decode()
    nv.nvrtcDestroyProgram(C.byref(prog))
    e.close()


def _emul_with_rhs(expr, consts):
    from helpers import ELLIPTIC, FN_BC, FN_JACOBIAN, FN_MAPPING, FN_RHS, FN_SOLUTION, MESH_REFINE_GLOBAL

    e = Emul(ELLIPTIC, MESH_REFINE_GLOBAL, 2, 12)
    e.set_function(FN_MAPPING, "0.9*y0 + 0.1*x0; 1.1*y1 - 0.05*x1", consts)
    e.set_function(FN_JACOBIAN, "0.9; 0; 0; 1.1", consts)
    e.set_function(FN_RHS, expr, consts)
    e.set_function(FN_BC, "0", consts)
    e.set_function(FN_SOLUTION, "0", consts)
    xy = np.array([[-0.5, 0.25], [0.3, 0.7]])
    e.setup(xy)
    return e, xy


def _generated_code_equals_interpreter(tmp_path, n, expr, consts):
    e, xy = _emul_with_rhs(expr, consts)
    src = _source(e)
    cpp = tmp_path / f"k{n}.cpp"
    cpp.write_text(SHIM + src + DRIVER)
    so = tmp_path / f"k{n}.so"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-Wno-unknown-pragmas", "-o", str(so), str(cpp)])
    lib = C.CDLL(str(so))
    N, ncells = len(xy), 4
    e.assemble(np.ones(N), np.ones(N))
    want = np.zeros(N * 18 * ncells)
    e.lib.emul_get_cell_scratch.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
    e.lib.emul_get_cell_scratch(e.h, dptr(want), want.size)
    slots, t0, t1 = C.c_void_p(), C.c_void_p(), C.c_void_p()
    dims = (C.c_int * 10)()
    e.lib.emul_table_pointers.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 3 + [C.POINTER(C.c_int)]
    e.lib.emul_table_pointers(e.h, C.byref(slots), C.byref(t0), C.byref(t1), dims)
    e.lib.emul_mesh_h.restype = C.c_double
    e.lib.emul_mesh_h.argtypes = [C.c_void_p]
    got = np.zeros_like(want)
    flag = C.c_int(0)
    lib.host_run.argtypes = [C.POINTER(C.c_int), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_double, C.c_void_p, C.POINTER(C.c_int)]
    lib.host_run(dims, e.lib.emul_mesh_h(e.h), slots, t0, t1, N, 1.0, dptr(got), C.byref(flag))
    got, want = got.reshape(N, 18, ncells), want.reshape(N, 18, ncells)
    assert np.isfinite(want[:, 10:14]).all(), expr
    assert np.array_equal(got[:, 10:14, :], want[:, 10:14, :]), expr
    e.close()


def test_every_opcode_is_lowered_like_the_interpreter_executes_it(tmp_path):
    """The expression list of test_expr_compile.py (all operators and functions of the grammar) as right-hand
    sides: the generated straight-line code, run on the host, must reproduce the interpreter bit for bit."""
    from test_expr_compile import CONSTS, EXPRS

    for n, expr in enumerate(EXPRS):
        _generated_code_equals_interpreter(tmp_path, n, expr, CONSTS)


def _random_expression(rng, depth):
    """A random member of the grammar exercised by tests/test_expr_random.py (all values finite on [-1,1]^4)."""
    if depth == 0 or rng.random() < 0.15:
        return rng.choice(["x0", "x1", "y0", "y1", "0.5", "2", "1.25", "7e-1", "D_2", "kappa_1"])
    a = _random_expression(rng, depth - 1)
    b = _random_expression(rng, depth - 1)
    return rng.choice([
        f"({a} + {b})", f"({a} - {b})", f"{a}*{b}", f"{a}/(2 + sin({b}))", f"-{a}", f"({a})^2", f"({a})^3",
        f"sin({a})", f"cos ({a})", f"tanh({a})", f"exp(sin({a}))", f"sqrt(abs({a}))", f"log(1 + abs({a}))",
        f"min({a}, {b})", f"max({a}, {b}, 0.5)", f"if({a} > {b}, {a}, -{b})", f"(({a} < {b}) | ({a} >= 1))",
        f"pow(1 + abs({a}), sin({b}))", f"atan2({a}, 2 + cos({b}))",
    ])


def test_generated_code_equals_interpreter_on_random_expressions(tmp_path):
    import random

    rng = random.Random(20261020)
    consts = {"D_2": 1.25, "kappa_1": 0.7}
    for n in range(16):
        _generated_code_equals_interpreter(tmp_path, 100 + n, _random_expression(rng, 4), consts)
