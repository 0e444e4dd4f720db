"""GPU parity of the elliptic micro path, through the C ABI, against the CPU oracle.

What is compared is what the reference's MicroSolver::run() produces for a given macro vector
(src/elliptic-elliptic/micro.cpp:140-182) and what MacroSolver::get_micro_bulk reads from it
(macro.cpp:134-157): system matrices after apply_boundary_values, right-hand sides, solutions,
CG iteration counts, bulk integrals.  Tolerance: 1e-10 relative L2 (BASELINE.json north_star).
"""
import os

import numpy as np
import pytest

from helpers import INPUT, OracleElliptic, assert_iterations_consistent, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-10

CASES = [
    ("elliptic_linear.prm", 1, 2), ("elliptic_linear.prm", 2, 3), ("elliptic_linear.prm", 3, 4),
    ("elliptic_non_linear.prm", 2, 3), ("elliptic_non_linear.prm", 3, 4),
    ("map_test.prm", 2, 3), ("arb_map_test.prm", 2, 3), ("rot_map_test.prm", 2, 3),
    ("simple_map_test.prm", 2, 3), ("lin_map_test.prm", 1, 2), ("scale_map_test.prm", 2, 3),
]


def _gpu_solver(prm, micro_ref, xy):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, prm)), 2 ** micro_ref)
    ms.set_grid_locations(xy)
    ms.setup()
    return ms


@pytest.mark.parametrize("prm,macro_ref,micro_ref", CASES)
def test_micro_run_matches_oracle(prm, macro_ref, micro_ref):
    o = OracleElliptic(prm, macro_ref, micro_ref)
    xy = o.grid_locations()
    u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
    o.micro_run(u)
    ms = _gpu_solver(prm, micro_ref, xy)
    ms.set_macro_solution(u)
    its = ms.run()
    x_gpu = ms.solutions()
    b_gpu = ms.righthandsides()
    for k in range(o.N):
        assert rel_l2(b_gpu[k], o.rhs(k)) < TOL
        assert rel_l2(x_gpu[k], o.solution(k)) < TOL
    for k in sorted({0, o.N // 2, o.N - 1}):
        A_o, A_g = o.matrix(k), ms.system_matrix(k)
        assert np.abs(A_o - A_g).max() <= 1e-12 * np.abs(A_o).max()
    # Dirichlet rows are satisfied from the start: the effective dimension is the interior node count
    assert_iterations_consistent(o, its, unknowns=(2 ** micro_ref - 1) ** 2)
    assert rel_l2(ms.coupling(), o.bulk()) < TOL
    np.testing.assert_allclose(ms.dof_points(), o.dof_points(), atol=1e-15)
    ms.close()
    o.close()


def test_repeated_runs_are_bitwise_reproducible():
    o = OracleElliptic("elliptic_non_linear.prm", 2, 3)
    xy = o.grid_locations()
    u = np.cos(xy[:, 0] + xy[:, 1])
    ms = _gpu_solver("elliptic_non_linear.prm", 3, xy)
    ms.set_macro_solution(u)
    ms.run()
    a = ms.solutions().copy()
    ms.run()
    b = ms.solutions()
    assert np.array_equal(a, b)
    ms.close()
    o.close()


def test_jacobi_preconditioner_gives_same_solution():
    from dealiiscale_b200 import capi

    o = OracleElliptic("arb_map_test.prm", 2, 3)
    xy = o.grid_locations()
    u = np.full(o.N, 5.0)
    o.micro_run(u)
    ms = _gpu_solver("arb_map_test.prm", 3, xy)
    ms.set_macro_solution(u)
    ms.precond = capi.PRECOND_JACOBI
    ms.run()
    x = ms.solutions()
    for k in range(o.N):
        assert rel_l2(x[k], o.solution(k)) < 1e-9
    ms.close()
    o.close()


def test_global_memory_cg_variant_matches(monkeypatch):
    monkeypatch.setenv("MSGPU_CG_VARIANT", "global")
    o = OracleElliptic("elliptic_linear.prm", 2, 3)
    xy = o.grid_locations()
    u = xy[:, 0] - xy[:, 1]
    o.micro_run(u)
    ms = _gpu_solver("elliptic_linear.prm", 3, xy)
    ms.set_macro_solution(u)
    ms.run()
    assert ms.stats()["cg_variant"] == 1
    x = ms.solutions()
    for k in range(o.N):
        assert rel_l2(x[k], o.solution(k)) < TOL
    ms.close()
    o.close()


def test_parse_error_is_reported():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver

    ms = MicroSolver(capi.ELLIPTIC, None, 4)
    with pytest.raises(capi.MsgpuError) as e:
        ms.set_function(capi.FN_RHS, "x0 + Abs(y0)")
    assert e.value.code == capi.ERR_PARSE
    with pytest.raises(capi.MsgpuError):
        ms.set_function(capi.FN_MAPPING, "y0")  # one component instead of two
    ms.close()
