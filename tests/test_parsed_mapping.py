"""N = 1 cross-check of the Robin / Neumann assembly on `input/parsed_mapping.prm` (BASELINE.json
configs[1]): the single pulled-back Robin system of the reference's `src/playground/transform.cpp:263-367`
is one micro system of the bio path (`src/bio-multiscale/micro.cpp:183-276`) with kappa_2 = kappa_4 = D_2 = 1,
kappa_1 = kappa_3 = 0, QGauss(8), variables `x, y` instead of `y0, y1` and the constant `pi`.  The checker is an
independent NumPy/SciPy restatement (`oracle/robin_single.py`), itself pinned analytically: its error against
`ref_solution` must fall with order 2, which is what transform.cpp prints."""
import os
import re
import sys

import numpy as np
import pytest

from helpers import (BIO, FN_BC_V1, FN_BC_V2, FN_BC_V3, FN_BC_V4, FN_JACOBIAN, FN_MAPPING, FN_RHS, FN_SOLUTION, INPUT,
                     MESH_REFINE_GLOBAL, ROOT, Emul, iptr, load_emul, rel_l2)

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import robin_single  # noqa: E402  (the checker)

# "defaults" = the entries transform.cpp declares (:102-116: rotation by pi/6 with consistent Robin / Neumann data)
PRMS = ["defaults", "parsed_mapping.prm", "ss_lin_equiv.prm", "ss_smt_equivalent.prm"]


Q = 8  # transform.cpp:264


def read(prm):
    return dict(robin_single.DEFAULTS) if prm == "defaults" else robin_single.read_prm(os.path.join(INPUT, prm))


def micro_names(expr):
    """transform.cpp uses FunctionParser's default variable names x, y for the micro coordinates."""
    return re.sub(r"\by\b", "y1", re.sub(r"\bx\b", "y0", expr))


def configure(obj, p):
    """Feed a transform.cpp parameter set into anything with set_function / set_scalars (Emul or MicroSolver)."""
    c = {"pi": np.pi}
    obj.set_function(FN_MAPPING, micro_names(p["mapping"]), c)
    obj.set_function(FN_JACOBIAN, micro_names(p["jac_mapping"]), c)
    # rhs and boundary data are evaluated at the MAPPED point (transform.cpp:295, :310-330) = the bio composition
    obj.set_function(FN_RHS, micro_names(p["rhs"]), c)
    obj.set_function(FN_SOLUTION, micro_names(p["solution"]), c)
    obj.set_function(FN_BC_V1, micro_names(p["left_robin"]), c)
    obj.set_function(FN_BC_V2, micro_names(p["right_robin"]), c)
    obj.set_function(FN_BC_V3, micro_names(p["up_neumann"]), c)
    obj.set_function(FN_BC_V4, micro_names(p["down_neumann"]), c)
    obj.set_scalars([0.0, 1.0, 0.0, 1.0, 1.0])  # kappa_1..4, D_2


def to_oracle_order(nc):
    """perm[reference dof] = oracle node (iy*m + ix), hyper_cube + refine_global numbering."""
    m = nc + 1
    r2i = np.zeros(m * m, dtype=np.int32)
    assert load_emul().emul_reference_numbering(MESH_REFINE_GLOBAL, nc, iptr(r2i)) == 0
    ix, iy = r2i // m, r2i % m
    return iy * m + ix


def test_checker_converges_to_the_manufactured_solution_with_order_two():
    """transform.cpp runs refinements 2..6 and tabulates the L2 error (transform.cpp:493-497).  Only the declared
    defaults are a consistent manufactured set: the committed parsed_mapping.prm / ss_*.prm leave the boundary data
    at zero or at the defaults of another map, so their error stagnates (2.19, 12.5, 4.08) — in the reference too."""
    errs = []
    for r in (2, 3, 4, 5):
        s = robin_single.RobinSystem(read("defaults"), 2 ** r, Q, Q).assemble()
        s.solve()
        errs.append(s.l2_error())
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert errs[-1] < 6e-4 and (rates > 1.95).all(), (errs, rates)


@pytest.mark.parametrize("prm", PRMS)
def test_kernel_bodies_assemble_the_transform_system(prm):
    """The product's assembly bodies (CPU harness), generic quadrature degree 8, general-F mode with faces."""
    nc = 8
    p = read(prm)
    s = robin_single.RobinSystem(p, nc, Q, Q).assemble()
    e = Emul(BIO, MESH_REFINE_GLOBAL, nc, Q)
    configure(e, p)
    e.setup(np.array([[0.0, 0.0]]))
    assert e.assemble(np.zeros(1), np.zeros(1)) == 0
    perm = to_oracle_order(nc)
    A = s.A.toarray()[np.ix_(perm, perm)]
    assert rel_l2(e.matrix(0), A) < 1e-13
    assert rel_l2(e.rhs(0), s.b[perm]) < 1e-13
    rc, its = e.solve(1e-10, 10000)  # transform.cpp:364
    assert rc == 0
    x = s.solve()[perm]
    assert rel_l2(e.solution(0), x) < 1e-8  # CG at 1e-10 absolute against the direct solve
    # the flux integrals the bio macro solver takes from this system (bio macro.cpp:254-296), with the same solution
    xo = np.zeros_like(x)
    xo[perm] = e.solution(0)
    cu, cw = e.coupling()
    for got, side in ((cu[0], 0), (cw[0], 1)):
        want = s.flux(side, 1.0, xo)
        assert abs(got - want) <= 1e-11 * max(1.0, abs(want))
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prm,refine", [("defaults", 5), ("parsed_mapping.prm", 4), ("parsed_mapping.prm", 6),
                                        ("ss_lin_equiv.prm", 5)])
def test_gpu_solves_the_transform_system(prm, refine):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver

    nc = 2 ** refine
    p = read(prm)
    s = robin_single.RobinSystem(p, nc, Q, Q).assemble()
    ms = MicroSolver(capi.BIO, None, nc, device=0, q_degree=Q, mesh_kind=capi.MESH_REFINE_GLOBAL)
    configure(ms, p)
    ms.set_grid_locations(np.array([[0.0, 0.0]]))
    ms.setup()
    ms.set_macro_solution(np.zeros(1), np.zeros(1))
    ms.assemble_system()
    pts = ms.dof_points()
    m = nc + 1
    ix = np.rint((pts[:, 0] + 1.0) / s.h).astype(int)
    iy = np.rint((pts[:, 1] + 1.0) / s.h).astype(int)
    perm = iy * m + ix
    assert rel_l2(ms.righthandsides()[0], s.b[perm]) < 1e-13
    if nc <= 16:
        assert rel_l2(ms.system_matrix(0), s.A.toarray()[np.ix_(perm, perm)]) < 1e-13
    ms.abs_tol = 1e-12
    its = ms.solve()
    x = ms.solutions()[0]
    xo = s.solve()[perm]
    assert rel_l2(x, xo) < 1e-10, its
    # the analytic pin, through the product: error against ref_solution as transform.cpp measures it
    xg = np.zeros_like(x)
    xg[perm] = x
    assert abs(s.l2_error(xg) - s.l2_error()) < 1e-9
    ms.close()
