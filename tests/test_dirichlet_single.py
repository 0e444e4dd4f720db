"""Second opinion on the elliptic micro path: `oracle/dirichlet_single.py`, a NumPy/SciPy restatement of
one Dirichlet micro system of `src/elliptic-elliptic/micro.cpp:108-171` that shares nothing with the C++ oracle, against
(i) its own analytic pin, (ii) the C++ oracle, (iii) the product's kernel bodies (CPU harness) and (iv) the GPU."""
import os
import sys

import numpy as np
import pytest

from helpers import ELLIPTIC, INPUT, MESH_REFINE_GLOBAL, ROOT, Emul, OracleElliptic, configure_elliptic, iptr, load_emul, rel_l2

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dirichlet_single  # noqa: E402  (the checker)

from dealiiscale_b200.prm import EllipticData  # noqa: E402

PRMS = ["elliptic_non_linear.prm", "rot_map_test.prm", "elliptic_linear.prm"]


def _strings(data):
    m = data.micro
    return {"mapping": m["mapping"], "jac_mapping": m["jac_mapping"], "micro_rhs": m["rhs"], "micro_bc": m["bc"],
            "micro_solution": m["solution"]}


def _perm(nc):
    m = nc + 1
    r2i = np.zeros(m * m, dtype=np.int32)
    assert load_emul().emul_reference_numbering(MESH_REFINE_GLOBAL, nc, iptr(r2i)) == 0
    return (r2i % m) * m + r2i // m


def _macro_value(data, x):
    import math

    return eval(data.params.get("macro_solution").replace("^", "**"),
                {"__builtins__": {}, "sin": math.sin, "cos": math.cos, "exp": math.exp, "x0": x[0], "x1": x[1]})


def test_checker_approaches_the_manufactured_micro_solution_with_order_two():
    data = EllipticData(os.path.join(INPUT, "elliptic_non_linear.prm"))
    x = (0.5, -0.5)
    errs = []
    for nc in (4, 8, 16):
        s = dirichlet_single.DirichletSystem(_strings(data), x, _macro_value(data, x), nc).assemble()
        s.solve()
        errs.append(s.l2_error())
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert (rates > 1.9).all(), (errs, rates)


@pytest.mark.parametrize("prm", PRMS)
def test_cpp_oracle_and_kernel_bodies_match_the_checker(prm):
    data = EllipticData(os.path.join(INPUT, prm))
    o = OracleElliptic(prm, 1, 3)        # 9 macro points, 8 x 8 micro cells
    xy = o.grid_locations()
    u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
    o.micro_run(u)
    e = Emul(ELLIPTIC, MESH_REFINE_GLOBAL, 8, 12)
    configure_elliptic(e, data)
    e.setup(xy)
    assert e.assemble(u) == 0
    e.solve(1e-12, 1000)
    perm = _perm(8)
    bulk_e, _ = e.coupling()
    l2_e, _ = e.errors()          # per-grid L2 error against micro_solution (micro.cpp:193-202)
    l2_o, _ = o.grid_errors()
    for k in (0, 4, 7):
        s = dirichlet_single.DirichletSystem(_strings(data), tuple(xy[k]), u[k], 8).assemble()
        x = s.solve()[perm]
        A = s.A[np.ix_(perm, perm)]
        assert np.abs(o.matrix(k) - A).max() <= 1e-12 * np.abs(A).max()
        assert rel_l2(o.rhs(k), s.b[perm]) < 1e-12
        assert rel_l2(o.solution(k), x) < 1e-10
        assert abs(o.bulk()[k] - s.bulk()) <= 1e-10 * max(1.0, abs(s.bulk()))
        assert np.abs(e.matrix(k) - A).max() <= 1e-12 * np.abs(A).max()
        assert rel_l2(e.rhs(k), s.b[perm]) < 1e-12
        assert rel_l2(e.solution(k), x) < 1e-10
        assert abs(bulk_e[k] - s.bulk()) <= 1e-10 * max(1.0, abs(s.bulk()))
        assert abs(l2_e[k] - s.l2_error()) <= 1e-9 and abs(l2_o[k] - s.l2_error()) <= 1e-9
    e.close()
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prm,nc", [("elliptic_non_linear.prm", 16), ("rot_map_test.prm", 8)])
def test_gpu_matches_the_checker(prm, nc):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver

    data = EllipticData(os.path.join(INPUT, prm))
    xy = np.array([[0.5, -0.5], [-1.0, 0.25], [0.0, 1.0]])
    u = np.array([0.3, -0.2, 1.1])
    ms = MicroSolver(capi.ELLIPTIC, data, nc)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_macro_solution(u)
    ms.run()
    x, c = ms.solutions(), ms.coupling()
    pts = ms.dof_points()
    h = 2.0 / nc
    perm = np.rint((pts[:, 1] + 1) / h).astype(int) * (nc + 1) + np.rint((pts[:, 0] + 1) / h).astype(int)
    for k in range(3):
        s = dirichlet_single.DirichletSystem(_strings(data), tuple(xy[k]), u[k], nc).assemble()
        assert rel_l2(x[k], s.solve()[perm]) < 1e-10
        assert abs(c[k] - s.bulk()) <= 1e-10 * max(1.0, abs(s.bulk()))
    ms.close()
