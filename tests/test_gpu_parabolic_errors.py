"""GPU parity, through the C ABI, of (i) the elliptic-parabolic micro solver RhoSolver
(src/elliptic-parabolic/rho_solver.cpp:55-237: L2 projection of init_rho, shared matrix
M + dt D L + Robin mass, per-grid right-hand sides, cold-started CG, PiSolver::get_micro_mass) and
(ii) the per-grid error / residual norms (micro.cpp:189-205, bio micro.cpp:367-378)."""
import os

import numpy as np
import pytest

from helpers import INPUT, OracleBio, OracleElliptic, OracleParabolic, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prm,h_inv,t_inv", [("linear_time.prm", 8, 4), ("linear_time.prm", 16, 16),
                                              ("full_nonlinear.prm", 11, 8), ("nonlinear-time.prm", 8, 4)])
def test_rho_solver_time_steps(prm, h_inv, t_inv):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import TwoPressureData

    o = OracleParabolic(prm, h_inv, h_inv, t_inv)
    xy = o.grid_locations()
    ms = MicroSolver(capi.PARABOLIC, TwoPressureData(os.path.join(INPUT, prm)), h_inv)
    ms.set_grid_locations(xy)
    ms.setup()
    x0 = ms.solutions()
    assert max(rel_l2(x0[k], o.solution(k)) for k in range(o.N)) < 1e-10      # projected initial data
    dt, t = 1.0 / t_inv, 0.0
    for step in range(3):
        t += dt
        pi = np.cos(xy[:, 0]) * xy[:, 1] + 2 + 0.1 * step
        o.micro_step(pi, dt, t)
        ms.set_macro_solution(pi)
        its = ms.iterate(dt, t)
        x, b = ms.solutions(), ms.righthandsides()
        for k in range(o.N):
            assert rel_l2(b[k], o.rhs(k)) < 1e-11
            assert rel_l2(x[k], o.solution(k)) < 1e-10
        assert np.abs(its - o.iters()).max() <= 1
        assert rel_l2(ms.coupling(), o.mass()) < 1e-11
        l2, h1 = ms.grid_errors()
        l2o, h1o = o.grid_errors()
        # H1 uses central differences with h = 1e-8 of the exact solution (AutoDerivativeFunction):
        # one ulp of difference between device and host sin/cos/exp becomes ~1e-8 in the gradient
        assert rel_l2(l2, l2o) < 1e-9 and rel_l2(h1, h1o) < 1e-6
    A = o.matrix()
    assert np.abs(A - ms.system_matrix(0)).max() <= 1e-13 * np.abs(A).max()
    ms.close()
    o.close()


def test_elliptic_grid_errors():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    for prm in ("elliptic_non_linear.prm", "rot_map_test.prm"):
        o = OracleElliptic(prm, 2, 3)
        xy = o.grid_locations()
        u = np.sin(xy[:, 0]) * xy[:, 1]
        o.micro_run(u)
        ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, prm)), 8)
        ms.set_grid_locations(xy)
        ms.setup()
        ms.set_macro_solution(u)
        ms.run()
        l2, h1 = ms.grid_errors()
        l2o, h1o = o.grid_errors()
        assert rel_l2(l2, l2o) < 1e-8 and rel_l2(h1, h1o) < 1e-6
        ms.close()
        o.close()


def test_bio_grid_errors_and_residuals():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    o = OracleBio("linear.prm", 2, 8)
    xy = o.grid_locations()
    u, w = np.sin(xy[:, 0]) * xy[:, 1], np.cos(xy[:, 1])
    ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, "linear.prm")), 8)
    ms.set_grid_locations(xy)
    ms.setup()
    for s in (1.0, 1.1):
        o.micro_run(s * u, w / s)
        ms.set_macro_solution(s * u, w / s)
        ms.run()
    l2, h1 = ms.grid_errors()
    l2o, h1o = o.grid_errors()
    assert rel_l2(l2, l2o) < 1e-8 and rel_l2(h1, h1o) < 1e-6
    assert rel_l2(ms.grid_residuals(), o.grid_residuals()) < 1e-6
    ms.close()
    o.close()
