"""SSOR-preconditioned CG (MSGPU_PRECOND_SSOR; BASELINE.json north_star names Jacobi / SSOR).  The reference itself
solves with PreconditionIdentity (micro.cpp:179), so there is no reference output to match: the checker is a NumPy
restatement of deal.II's preconditioned SolverCG recurrence with the same four-colour SSOR, on the matrices the GPU
assembled, plus the requirement that the preconditioned and the plain solve agree to the solver tolerance."""
import os

import numpy as np
import pytest

from helpers import INPUT, rel_l2

pytestmark = pytest.mark.gpu


def colour_ssor_pcg(A, b, x0, colour, omega, tol, max_it):
    """SolverCG::solve with a preconditioner (deal.II 9.1): g = A x - b; h = M^-1 g; d = -h; gh = g.h; loop."""
    n = len(b)
    D = np.diag(A).copy()
    lower = colour[None, :] < colour[:, None]
    upper = colour[None, :] > colour[:, None]
    L, U = np.where(lower, A, 0.0), np.where(upper, A, 0.0)
    order_f = [np.where(colour == c)[0] for c in range(4)]

    def apply(g):
        y = np.zeros(n)
        for idx in order_f:  # (D/w + L) y = g
            y[idx] = (g[idx] - omega * (L[idx] @ y)) / D[idx]
        z = y.copy()
        for idx in order_f[::-1]:  # (D/w + U) z = (D/w) y
            z[idx] = y[idx] - omega * (U[idx] @ z) / D[idx]
        return omega * (2.0 - omega) * z

    x = x0.copy()
    g = A @ x - b
    res = np.linalg.norm(g)
    if res <= tol:
        return x, 0
    h = apply(g)
    d = -h
    gh = g @ h
    it = 0
    while True:
        it += 1
        h = A @ d
        alpha = gh / (d @ h)
        x += alpha * d
        g += alpha * h
        res = np.linalg.norm(g)
        if res <= tol or it >= max_it:
            return x, it
        h = apply(g)
        beta = gh
        gh = g @ h
        beta = gh / beta
        d = beta * d - h


def _lattice_colour(pts, nc):
    h = 2.0 / nc
    ix = np.rint((pts[:, 0] + 1.0) / h).astype(int)
    iy = np.rint((pts[:, 1] + 1.0) / h).astype(int)
    return (ix & 1) | ((iy & 1) << 1)


@pytest.mark.parametrize("problem,prm,cells", [("elliptic", "elliptic_non_linear.prm", 8), ("bio", "biocase-b.prm", 9),
                                               ("elliptic", "rot_map_test.prm", 16), ("bio", "biocase-curved.prm", 20)])
def test_ssor_pcg_matches_numpy_restatement(problem, prm, cells):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData, EllipticData

    g = np.linspace(-0.8, 0.9, 2)
    xy = np.array([(a, b) for b in g for a in g])
    if problem == "elliptic":
        ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, prm)), cells)
    else:
        ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, prm)), cells)
    ms.set_grid_locations(xy)
    ms.setup()
    u, w = 1.0 + xy[:, 0], 0.5 - xy[:, 1]
    ms.set_macro_solution(u, w if problem == "bio" else None)
    ms.zero_solutions()
    ms.assemble_system()
    x_start = ms.solutions()           # elliptic: Dirichlet values in place (micro.cpp:165-169)
    its_id = ms.solve().copy()
    x_id = ms.solutions()
    colour = _lattice_colour(ms.dof_points(), cells)
    b = ms.righthandsides()
    ms.set_solutions(x_start)
    ms.precond = capi.PRECOND_SSOR
    its = ms.solve(want_residuals=True).copy()
    assert ms.stats()["cg_variant"] == 2
    x = ms.solutions()
    assert (ms.last_residuals <= 1e-12).all()
    for k in range(len(xy)):
        A = ms.system_matrix(k)
        xr, itr = colour_ssor_pcg(A, b[k], x_start[k], colour, 1.0, 1e-12, 10000)
        assert abs(int(its[k]) - itr) <= 2, (its[k], itr)
        assert rel_l2(x[k], xr) < 1e-9
        assert rel_l2(x[k], x_id[k]) < 1e-9
    assert its.sum() < 0.7 * its_id.sum()  # it is a preconditioner: clearly fewer steps than plain CG
    ms.close()


def test_ssor_at_the_baseline_size_and_its_limit():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    g = np.linspace(-1, 1, 3)
    xy = np.array([(a, b) for b in g for a in g])
    ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, "biocase.prm")), 64)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_macro_solution(1.0 + 0.5 * xy[:, 0], 2.0 - 0.25 * xy[:, 1])
    ms.assemble_system()
    its_id = ms.solve().copy()
    x_id = ms.solutions()
    ms.zero_solutions()
    ms.precond = capi.PRECOND_SSOR
    its = ms.solve().copy()
    x = ms.solutions()
    assert max(rel_l2(x[k], x_id[k]) for k in range(9)) < 1e-9
    assert its.max() < 0.6 * its_id.min()
    print("CG steps identity", its_id.mean(), "SSOR", its.mean())
    ms.close()
    big = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, "biocase.prm")), 80)  # 81 nodes per axis: more than one SM
    big.set_grid_locations(xy[:2])
    big.setup()
    big.set_macro_solution(np.ones(2), np.ones(2))
    big.assemble_system()
    big.precond = capi.PRECOND_SSOR
    with pytest.raises(capi.MsgpuError) as e:
        big.solve()
    assert e.value.code == capi.ERR_UNSUPPORTED
    big.close()
