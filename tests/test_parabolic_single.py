"""Second opinion on the elliptic-parabolic micro path: `oracle/parabolic_single.py` (NumPy/SciPy, shares nothing with
the C++ oracle) against its analytic pin, the C++ oracle, the product's kernel bodies (CPU harness) and the GPU.
Reference: src/elliptic-parabolic/rho_solver.cpp:55-229, pi_solver.cpp:178-197."""
import os
import sys

import numpy as np
import pytest

from helpers import INPUT, MESH_SUBDIVIDED, PARABOLIC, ROOT, Emul, OracleParabolic, configure_parabolic, iptr, load_emul, rel_l2

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import parabolic_single  # noqa: E402  (the checker)

from dealiiscale_b200.prm import TwoPressureData  # noqa: E402


def _strings(data):
    m = data.micro
    return {"micro_rhs": m["rhs"], "micro_bc_robin": m["bc_robin"], "micro_bc_neumann": m["bc_neumann"],
            "init_rho": m["init_rho"]}


def _perm(nc):
    m = nc + 1
    r2i = np.zeros(m * m, dtype=np.int32)
    assert load_emul().emul_reference_numbering(MESH_SUBDIVIDED, nc, iptr(r2i)) == 0
    return (r2i % m) * m + r2i // m


def test_checker_reproduces_the_stationary_linear_solution():
    data = TwoPressureData(os.path.join(INPUT, "linear_time.prm"))
    x = (0.25, -0.5)
    g = parabolic_single.RhoGrid(_strings(data), data.constants, x, 6)
    pts = g.points()
    exact = 3 * pts[:, 0] + pts[:, 1] + 2
    assert np.abs(g.rho - exact).max() < 1e-12
    dt = 0.125
    for step in range(1, 4):
        t = step * dt
        g.step(t / 2 + 2 * x[0] + 3 * x[1], dt, t)   # the exact macro value of linear_time.prm
        assert np.abs(g.rho - exact).max() < 1e-10
    assert abs(g.mass() - 8.0) < 1e-10                 # int (3 y0 + y1 + 2) over [-1,1]^2


@pytest.mark.parametrize("prm", ["linear_time.prm", "full_nonlinear.prm"])
def test_cpp_oracle_and_kernel_bodies_match_the_checker(prm):
    data = TwoPressureData(os.path.join(INPUT, prm))
    h_inv = 6
    o = OracleParabolic(prm, 2, h_inv, 8)
    xy = o.grid_locations()
    e = Emul(PARABOLIC, MESH_SUBDIVIDED, h_inv, 2)
    configure_parabolic(e, data)
    e.setup(xy)
    e.parabolic_init()
    perm = _perm(h_inv)
    grids = {k: parabolic_single.RhoGrid(_strings(data), data.constants, tuple(xy[k]), h_inv) for k in (0, 4, 8)}
    for k, g in grids.items():
        assert rel_l2(o.solution(k), g.rho[perm]) < 1e-10
        assert rel_l2(e.solution(k), g.rho[perm]) < 1e-10
    dt, t = 0.125, 0.0
    for step in range(2):
        t += dt
        pi = np.cos(xy[:, 0]) * xy[:, 1] + 2 + 0.1 * step
        o.micro_step(pi, dt, t)
        rc, _ = e.parabolic_step(pi, dt, t)
        assert rc == 0
        mass_e = e.mass_integrals()
        for k, g in grids.items():
            x = g.step(pi[k], dt, t)[perm]
            assert rel_l2(o.rhs(k), g.b[perm]) < 1e-12 and rel_l2(e.rhs(k), g.b[perm]) < 1e-12
            assert rel_l2(o.solution(k), x) < 1e-10 and rel_l2(e.solution(k), x) < 1e-10
            assert abs(o.mass()[k] - g.mass()) < 1e-10 * max(1, abs(g.mass()))
            assert abs(mass_e[k] - g.mass()) < 1e-10 * max(1, abs(g.mass()))
        A = grids[0].A[np.ix_(perm, perm)]
        assert np.abs(o.matrix() - A).max() <= 1e-12 * np.abs(A).max()
    e.close()
    o.close()


@pytest.mark.gpu
def test_gpu_matches_the_checker():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver

    prm, h_inv = "full_nonlinear.prm", 10
    data = TwoPressureData(os.path.join(INPUT, prm))
    xy = np.array([[0.5, -0.5], [-1.0, 0.25], [0.0, 1.0]])
    ms = MicroSolver(capi.PARABOLIC, data, h_inv)
    ms.set_grid_locations(xy)
    ms.setup()
    pts = ms.dof_points()
    h = 2.0 / h_inv
    perm = np.rint((pts[:, 1] + 1) / h).astype(int) * (h_inv + 1) + np.rint((pts[:, 0] + 1) / h).astype(int)
    grids = [parabolic_single.RhoGrid(_strings(data), data.constants, tuple(x), h_inv) for x in xy]
    x0 = ms.solutions()
    for k, g in enumerate(grids):
        assert rel_l2(x0[k], g.rho[perm]) < 1e-10
    dt, t = 0.0625, 0.0
    for step in range(3):
        t += dt
        pi = np.array([1.0 + 0.2 * step, -0.5, 2.0 * t])
        ms.set_macro_solution(pi)
        ms.iterate(dt, t)
        x, c = ms.solutions(), ms.coupling()
        for k, g in enumerate(grids):
            assert rel_l2(x[k], g.step(pi[k], dt, t)[perm]) < 1e-10
            assert abs(c[k] - g.mass()) < 1e-10 * max(1, abs(g.mass()))
    ms.close()
