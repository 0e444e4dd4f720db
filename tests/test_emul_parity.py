"""CPU tier: the product's per-thread kernel bodies (assembly, faces, gather, Dirichlet elimination,
right-hand sides, error / residual integrals, parabolic pieces), its expression compiler and its VM,
driven sequentially by tests/native and compared with the oracle.  The CG of this tier is a plain
host loop in the harness; the CUDA CG is covered by the -m gpu tests."""
import os

import numpy as np
import pytest

from helpers import (BIO, ELLIPTIC, INPUT, MESH_REFINE_GLOBAL, MESH_SUBDIVIDED, PARABOLIC, Emul, OracleBio,
                     OracleElliptic, OracleParabolic, configure_bio, configure_elliptic, configure_parabolic, rel_l2)
from dealiiscale_b200.prm import BioData, EllipticData, TwoPressureData

ELLIPTIC_SETS = ["elliptic_linear.prm", "elliptic_non_linear.prm", "arb_map_test.prm", "map_test.prm",
                 "rot_map_test.prm", "simple_map_test.prm", "lin_map_test.prm", "scale_map_test.prm"]


@pytest.mark.parametrize("vector_vm", [1, 0])
@pytest.mark.parametrize("prm", ELLIPTIC_SETS)
def test_elliptic_assembly_and_solution(prm, vector_vm):
    o = OracleElliptic(prm, 2, 3)
    e = Emul(ELLIPTIC, MESH_REFINE_GLOBAL, 8, 12)
    configure_elliptic(e, EllipticData(os.path.join(INPUT, prm)))
    e.set_vector_mode(vector_vm)
    xy = o.grid_locations()
    e.setup(xy)
    u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
    o.micro_run(u)
    assert e.assemble(u) == 0
    assert e.wrapped_nonzeros() == 0
    rc, its = e.solve(1e-12, 1000)
    assert rc == 0
    for k in range(o.N):
        A = o.matrix(k)
        assert np.abs(A - e.matrix(k)).max() <= 1e-13 * np.abs(A).max()
        assert rel_l2(e.rhs(k), o.rhs(k)) < 1e-13
        assert rel_l2(e.solution(k), o.solution(k)) < 1e-10
    assert rel_l2(e.coupling()[0], o.bulk()) < 1e-12
    l2, h1 = e.errors()
    l2o, h1o = o.grid_errors()
    assert rel_l2(l2, l2o) < 1e-9 and rel_l2(h1, h1o) < 1e-9
    e.close()
    o.close()


def test_moment_modes_are_chosen_from_the_expressions():
    expect = {"lin_map_test.prm": 1, "map_test.prm": 1, "arb_map_test.prm": 0, "elliptic_non_linear.prm": 0}
    for prm, mode in expect.items():
        e = Emul(ELLIPTIC, MESH_REFINE_GLOBAL, 2, 12)
        configure_elliptic(e, EllipticData(os.path.join(INPUT, prm)))
        e.setup(np.array([[0.1, 0.2]]))
        assert e.moment_mode() == mode, prm   # 0 general, 1 Jacobian constant in y
        e.close()


@pytest.mark.parametrize("prm", ["linear.prm", "biocase.prm", "biocase-a.prm", "biocase-b.prm", "biocase-c.prm",
                                 "biocase-curved.prm", "biocase-d.prm"])
def test_bio_assembly_solution_flux_and_residuals(prm):
    o = OracleBio(prm, 2, 8)
    data = BioData(os.path.join(INPUT, prm))
    e = Emul(BIO, MESH_SUBDIVIDED, 8, 12)
    configure_bio(e, data)
    xy = o.grid_locations()
    e.setup(xy)
    u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
    w = np.cos(xy[:, 1]) * xy[:, 0] - 0.1
    o.micro_run(u, w)
    assert e.assemble(u, w) == 0
    assert e.wrapped_nonzeros() == 0
    rc, _ = e.solve(1e-12, 10000)
    assert rc == 0
    tol = 1e-9 if prm in ("biocase-c.prm", "biocase-curved.prm") else 1e-10
    for k in range(o.N):
        A = o.matrix(k)
        assert np.abs(A - e.matrix(k)).max() <= 1e-13 * np.abs(A).max()
        assert rel_l2(e.rhs(k), o.rhs(k)) < 1e-13
        assert rel_l2(e.solution(k), o.solution(k)) < tol
    cu, cw = e.coupling()
    cu_o, cw_o = o.flux()
    assert rel_l2(cu, cu_o) < tol and rel_l2(cw, cw_o) < tol
    # second cycle: warm start + residual against the previous iterate
    o.micro_run(1.1 * u, 0.9 * w)
    e.assemble(1.1 * u, 0.9 * w)
    e.solve(1e-12, 10000)
    assert rel_l2(e.residuals(), o.grid_residuals()) < 1e-7
    if o.has_solution():
        l2, h1 = e.errors()
        l2o, h1o = o.grid_errors()
        assert rel_l2(l2, l2o) < 1e-9 and rel_l2(h1, h1o) < 1e-9
    e.close()
    o.close()


@pytest.mark.parametrize("prm", ["linear_time.prm", "full_nonlinear.prm"])
def test_parabolic_projection_time_steps_mass_and_errors(prm):
    h_inv, t_inv = 8, 8
    o = OracleParabolic(prm, h_inv, h_inv, t_inv)
    e = Emul(PARABOLIC, MESH_SUBDIVIDED, h_inv, 2)
    configure_parabolic(e, TwoPressureData(os.path.join(INPUT, prm)))
    xy = o.grid_locations()
    e.setup(xy)
    e.parabolic_init()
    assert max(rel_l2(e.solution(k), o.solution(k)) for k in range(o.N)) < 1e-10   # L2 projection of init_rho
    dt, t = 1.0 / t_inv, 0.0
    for step in range(3):
        t += dt
        pi = np.cos(xy[:, 0]) * xy[:, 1] + 2 + 0.1 * step
        o.micro_step(pi, dt, t)
        rc, its = e.parabolic_step(pi, dt, t)
        assert rc == 0
        for k in range(o.N):
            assert rel_l2(e.rhs(k), o.rhs(k)) < 1e-11
            assert rel_l2(e.solution(k), o.solution(k)) < 1e-10
        assert np.abs(its - o.iters()).max() <= 1
        assert rel_l2(e.mass_integrals(), o.mass()) < 1e-11
        l2, h1 = e.errors()
        l2o, h1o = o.grid_errors()
        assert rel_l2(l2, l2o) < 1e-9 and rel_l2(h1, h1o) < 1e-9
    A = o.matrix()
    assert np.abs(A - e.matrix(0)).max() <= 1e-13 * np.abs(A).max()
    e.close()
    o.close()


@pytest.mark.parametrize("macro_cells,micro_cells", [(1, 1), (1, 2), (2, 5), (1, 7)])
def test_bio_small_and_odd_meshes(macro_cells, micro_cells):
    o = OracleBio("biocase-b.prm", macro_cells, micro_cells)
    e = Emul(BIO, MESH_SUBDIVIDED, micro_cells, 12)
    configure_bio(e, BioData(os.path.join(INPUT, "biocase-b.prm")))
    xy = o.grid_locations()
    e.setup(xy)
    u, w = 1.0 + xy[:, 0], 0.5 - xy[:, 1]
    o.micro_run(u, w)
    assert e.assemble(u, w) == 0
    e.solve(1e-12, 10000)
    for k in range(o.N):
        assert rel_l2(e.solution(k), o.solution(k)) < 1e-10
    e.close()
    o.close()


def test_dof_numbering_matches_the_oracle():
    """The product's reference numbering (fe_tables.h) against the oracle's independent one
    (oracle/fem.h), for both mesh kinds; plus the first-touch pattern itself on a 2x2 mesh."""
    from helpers import iptr, load_emul, load_oracle

    lib, em = load_oracle(), load_emul()
    for kind, nc in ((0, 1), (0, 2), (0, 8), (0, 16), (1, 1), (1, 3), (1, 11), (1, 16)):
        m = nc + 1
        node_dof = np.zeros(m * m, dtype=np.int32)        # oracle: lexicographic node iy*m+ix -> dof
        assert lib.orc_numbering(kind, nc, iptr(node_dof)) == 0
        ref_to_internal = np.zeros(m * m, dtype=np.int32)  # product: dof -> ix*m+iy
        assert em.emul_reference_numbering(kind, nc, iptr(ref_to_internal)) == 0
        for dof, r in enumerate(ref_to_internal):
            ix, iy = divmod(int(r), m)
            assert node_dof[iy * m + ix] == dof
    node_dof = np.zeros(9, dtype=np.int32)
    lib.orc_numbering(1, 2, iptr(node_dof))
    assert node_dof.tolist() == [0, 1, 4, 2, 3, 5, 6, 7, 8]
    lib.orc_numbering(0, 2, iptr(node_dof))
    assert node_dof.tolist() == [0, 1, 4, 2, 3, 5, 6, 7, 8]
