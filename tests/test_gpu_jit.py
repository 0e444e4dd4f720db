"""GPU tier: the run-time specialised cell-moment kernel (NVRTC) against the bytecode interpreter.

Both evaluate the same bytecode with one IEEE operation per instruction, and the accumulation code of
the generated kernel mirrors body_cell_moments_vec statement by statement, so matrices and right-hand
sides must agree bit for bit (the expression values) or to the last few ulps (compiler contraction
of the accumulation into FMAs may differ between nvcc and NVRTC builds)."""
import os

import numpy as np
import pytest

from helpers import INPUT

pytestmark = pytest.mark.gpu


def _assembled(problem, prm, cells, xy, u, w, jit):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData, EllipticData

    old = os.environ.get("MSGPU_JIT")
    os.environ["MSGPU_JIT"] = "1" if jit else "0"
    try:
        data = (BioData if problem == "bio" else EllipticData)(os.path.join(INPUT, prm))
        ms = MicroSolver(capi.BIO if problem == "bio" else capi.ELLIPTIC, data, cells)
        ms.set_grid_locations(xy)
        ms.setup()
        st = ms.stats()
        assert st["jit_kernels"] == (1 if jit else 0), ms.jit_log()
        if problem == "bio":
            ms.set_macro_solution(u, w)
        else:
            ms.set_macro_solution(u)
        ms.assemble_system()
        mats = np.stack([ms.system_matrix(k) for k in range(len(xy))])
        rhs = ms.righthandsides()
        its = ms.solve()
        sol = ms.solutions()
        ms.close()
        return mats, rhs, its, sol
    finally:
        if old is None:
            os.environ.pop("MSGPU_JIT", None)
        else:
            os.environ["MSGPU_JIT"] = old


CASES = [
    ("bio", "biocase.prm", 8),            # const-F mode, identity map
    ("bio", "biocase-b.prm", 5),          # const-F mode, affine map, odd cell count
    ("bio", "biocase-curved.prm", 8),     # const-F mode, F depends on the macro position
    ("elliptic", "elliptic_linear.prm", 8),
    ("elliptic", "elliptic_non_linear.prm", 8),
    ("elliptic", "rot_map_test.prm", 4),
    ("elliptic", "arb_map_test.prm", 4),
]


@pytest.mark.parametrize("problem,prm,cells", CASES)
def test_specialised_kernel_matches_interpreter(problem, prm, cells):
    g = np.linspace(-0.9, 0.8, 3)
    xy = np.array([(a, b) for b in g for a in g])
    u, w = 1.0 + 0.5 * xy[:, 0], 0.25 - xy[:, 1]
    m_i, b_i, its_i, x_i = _assembled(problem, prm, cells, xy, u, w, jit=False)
    m_j, b_j, its_j, x_j = _assembled(problem, prm, cells, xy, u, w, jit=True)
    scale_m, scale_b = np.abs(m_i).max(), max(np.abs(b_i).max(), 1e-300)
    assert np.abs(m_j - m_i).max() <= 4e-16 * scale_m
    assert np.abs(b_j - b_i).max() <= 4e-15 * scale_b      # elliptic: b also collects A * (Dirichlet values)
    assert np.abs(x_j - x_i).max() <= 1e-11 * max(np.abs(x_i).max(), 1e-300)
    assert np.abs(its_j - its_i).max() <= 2


def test_jit_can_be_switched_off():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    os.environ["MSGPU_JIT"] = "0"
    try:
        ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, "biocase.prm")), 4)
        ms.set_grid_locations(np.zeros((1, 2)))
        ms.setup()
        assert ms.stats()["jit_kernels"] == 0
        assert "MSGPU_JIT=0" in ms.jit_log()
        ms.close()
    finally:
        os.environ.pop("MSGPU_JIT", None)


@pytest.mark.parametrize("problem,prm,cells", [("elliptic", "elliptic_non_linear.prm", 8), ("bio", "linear.prm", 8)])
def test_specialised_error_kernel_matches_interpreter(problem, prm, cells):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData, EllipticData

    g = np.linspace(-0.9, 0.8, 3)
    xy = np.array([(a, b) for b in g for a in g])
    u, w = 1.0 + 0.5 * xy[:, 0], 0.25 - xy[:, 1]
    out = {}
    for jit in (False, True):
        os.environ["MSGPU_JIT"] = "1" if jit else "0"
        try:
            data = (BioData if problem == "bio" else EllipticData)(os.path.join(INPUT, prm))
            ms = MicroSolver(capi.BIO if problem == "bio" else capi.ELLIPTIC, data, cells)
            ms.set_grid_locations(xy)
            ms.setup()
            ms.set_macro_solution(u, w if problem == "bio" else None)
            ms.run()
            out[jit] = ms.grid_errors()
            assert ms.stats()["jit_kernels"] == (2 if jit else 0), ms.jit_log()
            ms.close()
        finally:
            os.environ.pop("MSGPU_JIT", None)
    # same expression values (one IEEE operation per bytecode instruction in both); the surrounding
    # arithmetic may be contracted differently by nvcc and NVRTC
    np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-9)
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=1e-6)


@pytest.mark.parametrize("prm", ["linear_time.prm", "full_nonlinear.prm"])
def test_parabolic_time_steps_with_and_without_the_specialised_kernel(prm):
    """RhoSolver::assemble_system (rho_solver.cpp:134-147) integrates the source with QGauss(2) at every time step; the
    NVRTC kernel and the interpreter must give the same right-hand sides and the same steps."""
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import TwoPressureData

    g = np.linspace(-0.9, 0.8, 3)
    xy = np.array([(a, b) for b in g for a in g])
    pi = 1.0 + 0.5 * xy[:, 0] - 0.25 * xy[:, 1]
    out = {}
    old = os.environ.get("MSGPU_JIT")
    try:
        for jit in (False, True):
            os.environ["MSGPU_JIT"] = "1" if jit else "0"
            ms = MicroSolver(capi.PARABOLIC, TwoPressureData(os.path.join(INPUT, prm)), 16)
            ms.set_grid_locations(xy)
            ms.setup()
            assert ms.stats()["jit_kernels"] >= (1 if jit else 0), ms.jit_log()
            if not jit:
                assert ms.stats()["jit_kernels"] == 0
            ms.set_macro_solution(pi)
            its = [ms.iterate(0.05, 0.05 * (i + 1)).copy() for i in range(3)]
            out[jit] = (np.stack(its), ms.righthandsides(), ms.solutions())
            ms.close()
    finally:
        if old is None:
            os.environ.pop("MSGPU_JIT", None)
        else:
            os.environ["MSGPU_JIT"] = old
    (its_i, b_i, x_i), (its_j, b_j, x_j) = out[False], out[True]
    assert np.abs(b_j - b_i).max() <= 4e-15 * max(np.abs(b_i).max(), 1e-300)
    assert np.abs(x_j - x_i).max() <= 1e-12 * max(np.abs(x_i).max(), 1e-300)
    assert np.array_equal(its_i, its_j)
