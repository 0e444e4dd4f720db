"""GPU tier: edge cases of the C ABI — smallest meshes, a single grid, odd cell counts, error codes."""
import os

import numpy as np
import pytest

from helpers import INPUT, OracleBio, OracleElliptic, rel_l2

pytestmark = pytest.mark.gpu


def _bio(prm, cells, xy):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, prm)), cells)
    ms.set_grid_locations(xy)
    ms.setup()
    return ms


@pytest.mark.parametrize("macro_cells,micro_cells", [(1, 1), (1, 2), (2, 5), (1, 7), (3, 33)])
def test_bio_small_and_odd_meshes(macro_cells, micro_cells):
    o = OracleBio("biocase-b.prm", macro_cells, micro_cells)
    xy = o.grid_locations()
    u, w = 1.0 + xy[:, 0], 0.5 - xy[:, 1]
    o.micro_run(u, w)
    ms = _bio("biocase-b.prm", micro_cells, xy)
    ms.set_macro_solution(u, w)
    ms.run()
    x = ms.solutions()
    for k in range(o.N):
        assert rel_l2(x[k], o.solution(k)) < 1e-10
    cu, cw = ms.coupling()
    cu_o, cw_o = o.flux()
    assert rel_l2(cu, cu_o) < 1e-10 and rel_l2(cw, cw_o) < 1e-10
    ms.close()
    o.close()


def test_single_grid_elliptic():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    o = OracleElliptic("arb_map_test.prm", 1, 1)      # 9 macro DoFs, micro mesh with 2x2 cells (one interior node)
    xy = o.grid_locations()
    u = np.full(o.N, 5.0)
    o.micro_run(u)
    ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "arb_map_test.prm")), 2)
    ms.set_grid_locations(xy[4:5])                   # one grid only
    ms.setup()
    ms.set_macro_solution(u[4:5])
    its = ms.run()
    assert rel_l2(ms.solutions()[0], o.solution(4)) < 1e-12
    assert its[0] == o.iters()[4]
    ms.close()
    o.close()


def test_no_convergence_is_reported():
    from dealiiscale_b200 import capi

    o = OracleBio("biocase-curved.prm", 1, 16)
    ms = _bio("biocase-curved.prm", 16, o.grid_locations())
    ms.set_macro_solution(np.ones(o.N), np.ones(o.N))
    ms.max_it = 5
    with pytest.raises(capi.MsgpuError) as e:
        ms.run()
    assert e.value.code == capi.ERR_NO_CONVERGENCE     # SolverControl::NoConvergence
    assert (ms.last_iterations == 5).all()
    ms.close()
    o.close()


def test_non_positive_jacobian_is_reported():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver

    ms = MicroSolver(capi.ELLIPTIC, None, 4)
    ms.set_function(capi.FN_MAPPING, "y0*(1+x0);-y1")
    ms.set_function(capi.FN_JACOBIAN, "1+x0;0;0;-1")           # det F = -(1+x0) < 0
    ms.set_function(capi.FN_RHS, "1")
    ms.set_function(capi.FN_BC, "0")
    ms.set_grid_locations(np.array([[0.0, 0.0], [0.5, 0.5]]))
    ms.max_it = 50
    with pytest.raises(capi.MsgpuError) as e:
        ms.setup()
        ms.set_macro_solution(np.zeros(2))
        ms.run()
    assert e.value.code in (capi.ERR_JACOBIAN, capi.ERR_NO_CONVERGENCE)
    ms.close()


def test_call_order_errors():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "map_test.prm")), 4)
    with pytest.raises(capi.MsgpuError) as e:
        ms.setup()                                     # no grid locations yet
    assert e.value.code == capi.ERR_INVALID
    with pytest.raises(capi.MsgpuError):
        ms.assemble_system()                           # not set up
    missing = MicroSolver(capi.ELLIPTIC, None, 4)
    missing.set_grid_locations(np.zeros((1, 2)))
    with pytest.raises(capi.MsgpuError) as e:
        missing.setup()                                # no functions
    assert e.value.code == capi.ERR_INVALID
    ms.close()
    missing.close()


def test_full_size_properties_biocase_6_6():
    """BASELINE size (4225 DoFs per system; a 9-point macro sample): size-independent properties —
    the CG meets its stopping test, the solution is affine in the macro values, repeated runs are
    bit-identical, and the flux integrals are consistent with the solutions they are taken from."""
    g = np.linspace(-1, 1, 3)
    xy = np.array([(a, b) for b in g for a in g])
    ms = _bio("biocase.prm", 64, xy)
    u, w = 1.0 + 0.5 * xy[:, 0], 2.0 - 0.25 * xy[:, 1]
    sols, flux = [], []
    for s in (0.0, 1.0, 2.0, 1.0):
        ms.zero_solutions()
        ms.set_macro_solution(s * u, s * w)
        ms.assemble_system()
        ms.solve(want_residuals=True)
        assert (ms.last_residuals <= 1e-12).all() and ms.last_iterations.max() < 10000
        sols.append(ms.solutions())
        flux.append(np.array(ms.coupling()))
    assert np.array_equal(sols[1], sols[3]) and np.array_equal(flux[1], flux[3])
    assert rel_l2(sols[2] - sols[0], 2 * (sols[1] - sols[0])) < 1e-9
    assert rel_l2(flux[2] - flux[0], 2 * (flux[1] - flux[0])) < 1e-9
    A = ms.system_matrix(4)
    b = ms.righthandsides(4, 5)[0]
    assert np.linalg.norm(A @ sols[1][4] - b) < 1e-10      # the assembled system is the one that was solved
    assert np.allclose(A, A.T, rtol=0, atol=0)
    ms.close()


def test_assemble_begin_in_every_order():
    """msgpu_assemble_begin (include/msgpu.h) is optional and must be harmless wherever a host calls it: twice in a
    row, before inputs change (new scalars / a new set-up retire what is in flight), right before destruction, for the
    elliptic problem, and never for results: the solutions equal those of a plain assemble + solve bit for bit."""
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData, EllipticData

    g = np.linspace(-0.8, 0.9, 4)
    xy = np.array([(a, b) for b in g for a in g])
    u, w = 1.0 + 0.5 * xy[:, 0], 0.25 - xy[:, 1]

    def plain(scalars=None):
        ms = _bio("biocase-b.prm", 12, xy)
        if scalars is not None:
            ms.set_scalars(scalars)
        ms.set_macro_solution(u, w)
        its = ms.run()
        x, c = ms.solutions(), ms.coupling()
        ms.close()
        return x, its.copy(), c

    x0, its0, c0 = plain()
    ms = _bio("biocase-b.prm", 12, xy)
    ms.assemble_begin()
    ms.assemble_begin()                      # second call: nothing to add
    ms.set_macro_solution(u, w)
    its = ms.run()                           # joins what is in flight
    assert np.array_equal(ms.solutions(), x0) and np.array_equal(its, its0)
    cu, cw = ms.coupling()
    assert np.array_equal(cu, c0[0]) and np.array_equal(cw, c0[1])
    # new Robin coefficients while an assembly is in flight: it is retired, the next one sees the new values
    sc = list(BioData(os.path.join(INPUT, "biocase-b.prm")).scalars())  # kappa_1 .. kappa_4, D_2
    sc[1] *= 2.0
    sc[3] *= 0.5
    ms.assemble_begin()
    ms.set_scalars(sc)
    ms.zero_solutions()
    its = ms.run()
    x1, its1, _ = plain(sc)
    assert np.array_equal(ms.solutions(), x1) and np.array_equal(its, its1)
    assert not np.array_equal(x1, x0)
    ms.assemble_begin()
    ms.close()                               # destroyed with an assembly in flight

    o = OracleElliptic("elliptic_non_linear.prm", 1, 3)
    xe = o.grid_locations()
    ue = np.sin(xe[:, 0]) * xe[:, 1] + 0.3
    o.micro_run(ue)
    me = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "elliptic_non_linear.prm")), 8)
    me.set_grid_locations(xe)
    me.setup()
    me.assemble_begin()
    me.set_macro_solution(ue)
    me.run()
    xs = me.solutions()
    for k in range(o.N):
        assert rel_l2(xs[k], o.solution(k)) < 1e-10
    me.close()
    o.close()
