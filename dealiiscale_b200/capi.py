"""ctypes binding of include/msgpu.h — the stub a maintainer of a Python host would write.

The library must have been built (``python -m dealiiscale_b200.build`` or ``__graft_entry__.build()``);
there is NO fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_build", "libmsgpu.so")

OK, ERR_INVALID, ERR_PARSE, ERR_JACOBIAN, ERR_NO_CONVERGENCE, ERR_CUDA, ERR_COMM, ERR_UNSUPPORTED = range(8)
ELLIPTIC, BIO, PARABOLIC = 0, 1, 2
MESH_REFINE_GLOBAL, MESH_SUBDIVIDED = 0, 1
(FN_MAPPING, FN_JACOBIAN, FN_RHS, FN_BC, FN_SOLUTION, FN_BC_V1, FN_BC_V2, FN_BC_V3, FN_BC_V4, FN_INIT) = range(10)
PRECOND_IDENTITY, PRECOND_JACOBI, PRECOND_SSOR = 0, 1, 2

# every symbol include/msgpu.h declares
EXPORTED = [
    "msgpu_create", "msgpu_destroy", "msgpu_last_error", "msgpu_set_stream", "msgpu_synchronize",
    "msgpu_set_grid_locations", "msgpu_set_function", "msgpu_set_scalars", "msgpu_setup", "msgpu_num_grids",
    "msgpu_num_dofs", "msgpu_set_macro_solution", "msgpu_assemble", "msgpu_solve", "msgpu_time_step",
    "msgpu_coupling", "msgpu_errors", "msgpu_residuals", "msgpu_get_solutions", "msgpu_set_solutions",
    "msgpu_zero_solutions", "msgpu_get_rhs", "msgpu_get_dof_permutation", "msgpu_get_dof_points", "msgpu_get_matrix_dense",
    "msgpu_comm_unique_id", "msgpu_partition", "msgpu_comm_init", "msgpu_broadcast_macro", "msgpu_enable_timing",
    "msgpu_get_stats", "msgpu_jit_log", "msgpu_macro_setup", "msgpu_macro_set_system", "msgpu_macro_set_rhs",
    "msgpu_macro_set_solution", "msgpu_macro_get_solution", "msgpu_macro_solve", "msgpu_comm_p2p_export",
    "msgpu_comm_p2p_attach", "msgpu_set_exact_solution", "msgpu_comm_allgather", "msgpu_comm_p2p_detach",
    "msgpu_reload_switches", "msgpu_macro_set_product_load", "msgpu_norm_setup", "msgpu_norm_of_values",
    "msgpu_norm_of_residuals", "msgpu_norm_of_errors", "msgpu_assemble_begin",
]


class Desc(C.Structure):
    _fields_ = [("problem", C.c_int), ("mesh_kind", C.c_int), ("cells_per_axis", C.c_int), ("q_degree", C.c_int),
                ("device", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("ms_tables", C.c_double), ("ms_moments", C.c_double), ("ms_gather", C.c_double),
                ("ms_cg", C.c_double), ("ms_coupling", C.c_double), ("launches", C.c_longlong),
                ("cg_iterations_total", C.c_longlong), ("cg_variant", C.c_int), ("jit_kernels", C.c_int)]


class MsgpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"msgpu error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load libmsgpu.so.  torch is imported first so that its bundled NCCL is the one both share."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m dealiiscale_b200.build`; there is no CPU fallback")
    try:
        import torch  # noqa: F401  (plumbing only: device selection, streams, torch.distributed)
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    lib.msgpu_create.argtypes = [C.POINTER(vp), C.POINTER(Desc)]
    lib.msgpu_destroy.argtypes = [vp]
    lib.msgpu_destroy.restype = None
    lib.msgpu_last_error.argtypes = [vp]
    lib.msgpu_last_error.restype = C.c_char_p
    lib.msgpu_jit_log.argtypes = [vp]
    lib.msgpu_jit_log.restype = C.c_char_p
    lib.msgpu_set_stream.argtypes = [vp, vp]
    lib.msgpu_synchronize.argtypes = [vp]
    lib.msgpu_set_grid_locations.argtypes = [vp, dp, C.c_int]
    lib.msgpu_set_function.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(C.c_char_p), dp, C.c_int, C.c_int]
    lib.msgpu_set_scalars.argtypes = [vp, dp, C.c_int]
    lib.msgpu_setup.argtypes = [vp]
    lib.msgpu_num_grids.argtypes = [vp]
    lib.msgpu_num_dofs.argtypes = [vp]
    lib.msgpu_set_macro_solution.argtypes = [vp, dp, dp]
    lib.msgpu_assemble.argtypes = [vp]
    lib.msgpu_assemble_begin.argtypes = [vp]
    lib.msgpu_solve.argtypes = [vp, C.c_double, C.c_int, C.c_int, ip, dp]
    lib.msgpu_time_step.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, ip]
    lib.msgpu_coupling.argtypes = [vp, dp, dp]
    lib.msgpu_errors.argtypes = [vp, dp, dp]
    lib.msgpu_residuals.argtypes = [vp, dp]
    lib.msgpu_set_exact_solution.argtypes = [vp, C.c_double]
    lib.msgpu_comm_allgather.argtypes = [vp, dp, dp, C.c_int]
    lib.msgpu_comm_p2p_detach.argtypes = [vp]
    lib.msgpu_get_solutions.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.msgpu_set_solutions.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.msgpu_zero_solutions.argtypes = [vp]
    lib.msgpu_get_rhs.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.msgpu_get_dof_permutation.argtypes = [vp, ip]
    lib.msgpu_get_dof_points.argtypes = [vp, dp]
    lib.msgpu_get_matrix_dense.argtypes = [vp, C.c_int, dp]
    lib.msgpu_comm_unique_id.argtypes = [C.c_char_p]
    lib.msgpu_partition.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
    lib.msgpu_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.msgpu_broadcast_macro.argtypes = [vp, dp, dp, C.c_int]
    lib.msgpu_enable_timing.argtypes = [vp, C.c_int]
    lib.msgpu_get_stats.argtypes = [vp, C.POINTER(Stats), C.c_int]
    _lib = lib
    return lib


def partition(n_total: int, n_ranks: int, rank: int):
    """Pure-Python twin of msgpu_partition (used by the host logic and the gloo tests)."""
    block = (n_total + n_ranks - 1) // n_ranks
    return min(rank * block, n_total), min((rank + 1) * block, n_total)
