"""Python host-side mirror of the reference's micro solvers on top of the C ABI.

``MicroSolver`` keeps the method names of the reference classes so that the parity tests read like
the reference's own driver code:

=====================================  =======================================================
reference (C++)                        here
=====================================  =======================================================
``MicroSolver(data, refine_level)``    ``MicroSolver(problem, data, cells_per_axis, ...)``
``set_grid_locations(locations)``      ``set_grid_locations(xy)``            micro.cpp:242-245
``setup()``                            ``setup()``                           micro.cpp:32-37
``set_macro_solution(&sol, &dofh)``    ``set_macro_solution(u, w=None)``     micro.cpp:97-100
``run()`` / ``assemble_and_solve_all``  ``run()``                             micro.cpp:218-221
``solutions``                          ``solutions()``                       micro.h:97
``set_exact_solution()``               ``set_exact_solution(t=0)``           micro.cpp:224-234
``RhoSolver::iterate(dt)``             ``iterate(dt, t)``                    rho_solver.cpp:232-237
``MacroSolver::compute_microscopic_contribution``  ``coupling()``           macro.cpp:162-166
=====================================  =======================================================

All numerical work happens in libmsgpu.so on the GPU; this module only marshals arguments.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .prm import BioData, EllipticData, TwoPressureData

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class MicroSolver:
    # SolverControl settings of the three drivers (micro.cpp:176, bio micro.cpp:310, rho_solver.cpp:223)
    DEFAULT_MAX_IT = {capi.ELLIPTIC: 1000, capi.BIO: 10000, capi.PARABOLIC: 10000}
    DEFAULT_Q = {capi.ELLIPTIC: 12, capi.BIO: 12, capi.PARABOLIC: 2}
    DEFAULT_MESH = {capi.ELLIPTIC: capi.MESH_REFINE_GLOBAL, capi.BIO: capi.MESH_SUBDIVIDED,
                    capi.PARABOLIC: capi.MESH_SUBDIVIDED}

    def __init__(self, problem: int, data, cells_per_axis: int, device: int = -1, q_degree: int | None = None,
                 mesh_kind: int | None = None):
        self.lib = capi.load()
        self.problem = problem
        self.data = data
        desc = capi.Desc(problem, self.DEFAULT_MESH[problem] if mesh_kind is None else mesh_kind, cells_per_axis,
                         self.DEFAULT_Q[problem] if q_degree is None else q_degree, device)
        h = C.c_void_p()
        rc = self.lib.msgpu_create(C.byref(h), C.byref(desc))
        if rc != capi.OK:
            raise capi.MsgpuError(rc, "msgpu_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.n_dofs = (cells_per_axis + 1) ** 2
        self.num_grids = 0
        self.n_total = 0
        self.k_offset = 0
        self.abs_tol = 1e-12
        self.max_it = self.DEFAULT_MAX_IT[problem]
        self.precond = capi.PRECOND_IDENTITY
        self.last_iterations = None
        if data is not None:
            self._configure(data)

    # ---- plumbing -------------------------------------------------------------------------
    def _check(self, rc):
        if rc != capi.OK:
            raise capi.MsgpuError(rc, self.lib.msgpu_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.msgpu_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_function(self, slot, expr, constants=None, time_dependent=False):
        constants = constants or {}
        names = [k.encode() for k in constants]
        arr = (C.c_char_p * max(1, len(names)))(*names)
        vals = np.array(list(constants.values()) or [0.0], dtype=np.float64)
        self._check(self.lib.msgpu_set_function(self.h, slot, expr.encode(), arr, _d(vals), len(names),
                                                int(time_dependent)))

    def set_scalars(self, values):
        v = np.asarray(values, dtype=np.float64)
        self._check(self.lib.msgpu_set_scalars(self.h, _d(v), len(v)))

    def _configure(self, data):
        c = data.constants
        m = data.micro
        if isinstance(data, EllipticData):
            self.set_function(capi.FN_MAPPING, m["mapping"], c)
            self.set_function(capi.FN_JACOBIAN, m["jac_mapping"], c)
            self.set_function(capi.FN_RHS, m["rhs"], c)
            self.set_function(capi.FN_BC, m["bc"], c)
            self.set_function(capi.FN_SOLUTION, m["solution"], c)
        elif isinstance(data, BioData):
            self.set_function(capi.FN_MAPPING, m["mapping"], c)
            self.set_function(capi.FN_JACOBIAN, m["jac_mapping"], c)
            self.set_function(capi.FN_RHS, m["rhs"], c)
            self.set_function(capi.FN_SOLUTION, m["solution"], c)
            for i, slot in enumerate((capi.FN_BC_V1, capi.FN_BC_V2, capi.FN_BC_V3, capi.FN_BC_V4), 1):
                self.set_function(slot, m[f"bc_v_{i}"], c)
            self.set_scalars(data.scalars())
        elif isinstance(data, TwoPressureData):
            self.set_function(capi.FN_RHS, m["rhs"], c, True)
            self.set_function(capi.FN_SOLUTION, m["solution"], c, True)
            self.set_function(capi.FN_BC_V1, m["bc_robin"], c, True)
            self.set_function(capi.FN_BC_V2, m["bc_neumann"], c, True)
            self.set_function(capi.FN_INIT, m["init_rho"], c, False)
            self.set_scalars(data.scalars())
        else:
            raise TypeError("unknown data class")

    def use_stream(self, cuda_stream: int):
        """Run on an existing stream (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        self._check(self.lib.msgpu_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.msgpu_synchronize(self.h))

    def reload_switches(self):
        """Read the MSGPU_* A/B switches from the environment again (they are read once, when the solver is created)."""
        self._check(self.lib.msgpu_reload_switches(self.h))

    # ---- the reference interface ------------------------------------------------------------
    def set_grid_locations(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        self.num_grids = xy.shape[0]
        if not self.n_total:
            self.n_total = self.num_grids
        self._check(self.lib.msgpu_set_grid_locations(self.h, _d(xy), self.num_grids))

    def get_num_grids(self):
        return self.num_grids

    def setup(self):
        self._check(self.lib.msgpu_setup(self.h))

    def set_macro_solution(self, u, w=None):
        u = np.ascontiguousarray(u, dtype=np.float64)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float64)
            wp = _d(w)
        self._check(self.lib.msgpu_set_macro_solution(self.h, _d(u), wp))

    def assemble_system(self):
        self._check(self.lib.msgpu_assemble(self.h))

    def assemble_begin(self):
        """bio: the macro-independent part of the next assemble_system() on a second stream (msgpu_assemble_begin)."""
        self._check(self.lib.msgpu_assemble_begin(self.h))

    def solve(self, want_residuals=False):
        its = np.zeros(self.num_grids, dtype=np.int32)
        res = np.zeros(self.num_grids) if want_residuals else None
        rc = self.lib.msgpu_solve(self.h, self.abs_tol, self.max_it, self.precond, _i(its),
                                  _d(res) if want_residuals else None)
        self.last_iterations = its
        self.last_residuals = res
        self._check(rc)
        return its

    def run(self):
        """assemble_system(); solve();  (micro.cpp:218-221 / bio micro.cpp:292-305)"""
        self.assemble_system()
        return self.solve()

    assemble_and_solve_all = run

    def iterate(self, dt, t):
        its = np.zeros(self.num_grids, dtype=np.int32)
        rc = self.lib.msgpu_time_step(self.h, dt, t, self.abs_tol, self.max_it, _i(its))
        self.last_iterations = its
        self._check(rc)
        return its

    def coupling(self):
        """The micro->macro integrals of all grids (allgathered when a communicator is attached)."""
        ntot = self.n_total or self.num_grids
        c0 = np.zeros(ntot)
        c1 = np.zeros(ntot)
        self._check(self.lib.msgpu_coupling(self.h, _d(c0), _d(c1)))
        return (c0, c1) if self.problem == capi.BIO else c0

    def grid_errors(self):
        l2 = np.zeros(self.num_grids)
        h1 = np.zeros(self.num_grids)
        self._check(self.lib.msgpu_errors(self.h, _d(l2), _d(h1)))
        return l2, h1

    def grid_residuals(self):
        l2 = np.zeros(self.num_grids)
        self._check(self.lib.msgpu_residuals(self.h, _d(l2)))
        return l2

    def set_exact_solution(self, t=0.0):
        """solutions[k] = L2 projection of the exact micro solution at x_k (micro.cpp:224-234, bio micro.cpp:381-391,
        rho_solver.cpp:353-359; ``t`` only matters for the parabolic problem)."""
        self._check(self.lib.msgpu_set_exact_solution(self.h, float(t)))

    def solutions(self, k0=0, k1=None):
        k1 = self.num_grids if k1 is None else k1
        out = np.zeros((k1 - k0, self.n_dofs))
        self._check(self.lib.msgpu_get_solutions(self.h, k0, k1, _d(out)))
        return out

    def set_solutions(self, values, k0=0):
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, self.n_dofs)
        self._check(self.lib.msgpu_set_solutions(self.h, k0, k0 + values.shape[0], _d(values)))

    def zero_solutions(self):
        self._check(self.lib.msgpu_zero_solutions(self.h))

    def coupling_on_device(self):
        """Coupling integrals (+ allgather) without the copy to the host."""
        self._check(self.lib.msgpu_coupling(self.h, None, None))

    def righthandsides(self, k0=0, k1=None):
        k1 = self.num_grids if k1 is None else k1
        out = np.zeros((k1 - k0, self.n_dofs))
        self._check(self.lib.msgpu_get_rhs(self.h, k0, k1, _d(out)))
        return out

    def system_matrix(self, k):
        out = np.zeros((self.n_dofs, self.n_dofs))
        self._check(self.lib.msgpu_get_matrix_dense(self.h, k, _d(out)))
        return out

    def dof_points(self):
        out = np.zeros((self.n_dofs, 2))
        self._check(self.lib.msgpu_get_dof_points(self.h, _d(out)))
        return out

    def dof_permutation(self):
        out = np.zeros(self.n_dofs, dtype=np.int32)
        self._check(self.lib.msgpu_get_dof_permutation(self.h, _i(out)))
        return out

    # ---- multi-GPU ----------------------------------------------------------------------------
    def attach_communicator(self, unique_id: bytes, rank: int, n_ranks: int, n_total: int):
        k0, k1 = capi.partition(n_total, n_ranks, rank)
        self.n_total = n_total
        self.k_offset = k0
        self._check(self.lib.msgpu_comm_init(self.h, unique_id, rank, n_ranks, n_total, k0))

    def p2p_export(self) -> bytes:
        """Handle of this rank's exchange buffer (to be gathered from all ranks by the host)."""
        buf = C.create_string_buffer(64)
        self._check(self.lib.msgpu_comm_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach(self, handles):
        """handles: one 64-byte handle per rank, in rank order.  coupling() then exchanges through peer memory."""
        blob = b"".join(handles)
        self._check(self.lib.msgpu_comm_p2p_attach(self.h, blob))

    def allgather(self, *local):
        """Per-grid scalars of all ranks (1 or 2 arrays of this rank's grids -> arrays of n_total entries), e.g. the
        error norms the stop test aggregates; ``msgpu_comm_allgather``."""
        fields = len(local)
        loc = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in local]))
        out = np.zeros((fields, self.n_total or self.num_grids))
        self._check(self.lib.msgpu_comm_allgather(self.h, _d(loc), _d(out), fields))
        return out[0] if fields == 1 else tuple(out)

    def p2p_detach(self):
        """Back to the NCCL allgather (e.g. after ``p2p_attach`` failed on some rank)."""
        self._check(self.lib.msgpu_comm_p2p_detach(self.h))

    def broadcast_macro(self, u_all, w_all=None, root=0):
        u_all = np.ascontiguousarray(u_all, dtype=np.float64)
        wp = None
        if w_all is not None:
            w_all = np.ascontiguousarray(w_all, dtype=np.float64)
            wp = _d(w_all)
        self._check(self.lib.msgpu_broadcast_macro(self.h, _d(u_all), wp, root))
        return u_all, w_all

    # ---- measurement ----------------------------------------------------------------------------
    def enable_timing(self, on=True):
        self._check(self.lib.msgpu_enable_timing(self.h, int(on)))

    def stats(self, reset=False):
        s = capi.Stats()
        self._check(self.lib.msgpu_get_stats(self.h, C.byref(s), int(reset)))
        return {f: getattr(s, f) for f, _ in capi.Stats._fields_}

    # ---- the macro system(s) on the device (include/msgpu.h, msgpu_macro_*) ----
    def macro_setup(self, n_fields, mesh_kind, cells_per_axis):
        self._check(self.lib.msgpu_macro_setup(self.h, int(n_fields), int(mesh_kind), int(cells_per_axis)))
        self._macro_n = (cells_per_axis + 1) ** 2
        self._macro_fields = n_fields

    def macro_set_system(self, field, source, rowstart, col, A, C=None, b0=None, x0=None):
        rs = np.ascontiguousarray(rowstart, dtype=np.int32)
        cl = np.ascontiguousarray(col, dtype=np.int32)
        keep = [np.ascontiguousarray(v, dtype=np.float64) if v is not None else None for v in (A, C, b0, x0)]
        ptr = [_d(v) if v is not None else None for v in keep]
        self._check(self.lib.msgpu_macro_set_system(self.h, int(field), int(source), _i(rs), _i(cl), *ptr))

    def macro_set_rhs(self, field, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        self._check(self.lib.msgpu_macro_set_rhs(self.h, int(field), _d(b)))

    def macro_set_solution(self, field, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._check(self.lib.msgpu_macro_set_solution(self.h, int(field), _d(x)))

    def macro_get_solution(self, field):
        out = np.zeros(self._macro_n)
        self._check(self.lib.msgpu_macro_get_solution(self.h, int(field), _d(out)))
        return out

    def macro_solve(self, tol=1e-12, max_it=10000, adopt=True):
        its = np.zeros(self._macro_fields, dtype=np.int32)
        self._check(self.lib.msgpu_macro_solve(self.h, C.c_double(tol), int(max_it), int(adopt), _i(its)))
        return its

    def jit_log(self) -> str:
        """Why the run-time specialised kernels are not in use ('' when they are)."""
        return self.lib.msgpu_jit_log(self.h).decode()


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = capi.load().msgpu_comm_unique_id(buf)
    if rc != capi.OK:
        raise capi.MsgpuError(rc, "ncclGetUniqueId failed")
    return buf.raw
