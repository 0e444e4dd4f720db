"""GPU parity of MicroSolver::set_exact_solution (src/elliptic-elliptic/micro.cpp:224-234,
src/bio-multiscale/micro.cpp:381-391, src/elliptic-parabolic/rho_solver.cpp:353-359): every solution vector
becomes the L2 projection (VectorTools::project) of the exact micro solution at that grid's macro point.  The
reference uses it for its micro-exact / micro-error dumps (manager.cpp:107-128)."""
import os

import numpy as np
import pytest

from helpers import INPUT, OracleBio, OracleElliptic, OracleParabolic, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prm,macro_ref,micro_ref", [("elliptic_non_linear.prm", 2, 3), ("rot_map_test.prm", 1, 4),
                                                      ("map_test.prm", 2, 2)])
def test_elliptic_exact_solution_projection(prm, macro_ref, micro_ref):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    o = OracleElliptic(prm, macro_ref, micro_ref)
    xy = o.grid_locations()
    o.set_exact_solution()
    ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, prm)), 2 ** micro_ref)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_exact_solution()
    x = ms.solutions()
    assert max(rel_l2(x[k], o.solution(k)) for k in range(o.N)) < 1e-10
    # the projection of the exact solution has (almost) no L2 error against it: far below the discretisation
    # error of the computed solution
    l2, _ = ms.grid_errors()
    u = np.sin(xy[:, 0]) * xy[:, 1]
    ms.set_macro_solution(u)
    ms.run()  # the context is still usable: assembly rebuilds the right-hand sides the projection overwrote
    o.micro_run(u)
    x = ms.solutions()
    assert max(rel_l2(x[k], o.solution(k)) for k in range(o.N)) < 1e-10
    assert np.isfinite(l2).all()
    ms.close()
    o.close()


@pytest.mark.parametrize("prm", ["linear.prm", "biocase-curved.prm"])
def test_bio_exact_solution_projection(prm):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    o = OracleBio(prm, 2, 8)
    xy = o.grid_locations()
    o.set_exact_solution()
    ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, prm)), 8)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_exact_solution()
    x = ms.solutions()
    assert max(rel_l2(x[k], o.solution(k)) for k in range(o.N)) < 1e-10
    ms.close()
    o.close()


def test_parabolic_exact_solution_projection():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import TwoPressureData

    prm, h_inv = "linear_time.prm", 8
    o = OracleParabolic(prm, h_inv, h_inv, 4)
    xy = o.grid_locations()
    ms = MicroSolver(capi.PARABOLIC, TwoPressureData(os.path.join(INPUT, prm)), h_inv)
    ms.set_grid_locations(xy)
    ms.setup()
    for t in (0.0, 0.25):
        o.set_exact_solution(t)
        ms.set_exact_solution(t)
        x = ms.solutions()
        assert max(rel_l2(x[k], o.solution(k)) for k in range(o.N)) < 1e-10
    ms.close()
    o.close()
