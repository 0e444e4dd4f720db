"""GPU tier: the macro system(s) on the device (msgpu_macro_*).

A small macro problem (Laplace + mass on the mesh of the grid locations, Dirichlet rows eliminated the
way MatrixTools::apply_boundary_values does) is built with NumPy; the device result of
b = b0 + C c -> CG -> adopt must equal the dense solve with the coupling vector the device itself produced."""
import os

import numpy as np
import pytest

from helpers import INPUT

pytestmark = pytest.mark.gpu


def _q1_1d(nc):
    h = 2.0 / nc
    K = np.zeros((nc + 1, nc + 1))
    M = np.zeros((nc + 1, nc + 1))
    for c in range(nc):
        K[c:c + 2, c:c + 2] += np.array([[1, -1], [-1, 1]]) / h
        M[c:c + 2, c:c + 2] += np.array([[2, 1], [1, 2]]) * h / 6
    return K, M


def _macro_problem(points, nc, dirichlet_all, rng):
    """A, C, b0, x0 in the numbering of `points` (rows: x, y of every DoF)."""
    m = nc + 1
    ix = np.rint((points[:, 0] + 1) / 2 * nc).astype(int)
    iy = np.rint((points[:, 1] + 1) / 2 * nc).astype(int)
    lex = iy * m + ix
    K1, M1 = _q1_1d(nc)
    A_lex = np.kron(M1, K1) + np.kron(K1, M1) + 0.7 * np.kron(M1, M1)
    C_lex = -np.kron(M1, M1)
    A = A_lex[np.ix_(lex, lex)].copy()
    C = C_lex[np.ix_(lex, lex)].copy()
    pattern = A != 0
    b0 = rng.standard_normal(m * m)
    x0 = np.zeros(m * m)
    on_b = (ix == 0) | (ix == nc) if not dirichlet_all else (ix == 0) | (ix == nc) | (iy == 0) | (iy == nc)
    for d in np.nonzero(on_b)[0]:                 # apply_boundary_values with column elimination
        v = 0.3 + 0.1 * points[d, 0] - 0.2 * points[d, 1]
        a_dd = A[d, d]
        A[d, :] = 0.0
        A[d, d] = a_dd
        b0[d] = v * a_dd
        rows = np.nonzero(A[:, d])[0]
        for r in rows:
            if r != d:
                b0[r] -= A[r, d] * v
                A[r, d] = 0.0
        C[d, :] = 0.0
        x0[d] = v
    rowstart = np.zeros(m * m + 1, dtype=np.int32)
    cols, av, cv = [], [], []
    for i in range(m * m):
        js = np.nonzero(pattern[i])[0]
        js = np.concatenate(([i], js[js != i]))     # deal.II: diagonal first
        cols.extend(js)
        av.extend(A[i, js])
        cv.extend(C[i, js])
        rowstart[i + 1] = len(cols)
    return A, C, b0, x0, rowstart, np.array(cols, dtype=np.int32), np.array(av), np.array(cv)


@pytest.mark.parametrize("problem,prm,nc", [("elliptic", "elliptic_linear.prm", 8), ("bio", "biocase-b.prm", 6),
                                            ("bio", "biocase.prm", 1)])
def test_device_macro_solve_matches_dense_solve(problem, prm, nc):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData, EllipticData

    rng = np.random.default_rng(7)
    bio = problem == "bio"
    kind = capi.MESH_SUBDIVIDED if bio else capi.MESH_REFINE_GLOBAL
    data = (BioData if bio else EllipticData)(os.path.join(INPUT, prm))
    # grid locations = DoF points of a macro mesh of the same kind and size as the micro mesh
    ms = MicroSolver(capi.BIO if bio else capi.ELLIPTIC, data, nc)
    ms.set_grid_locations(np.zeros((1, 2)))
    ms.setup()
    pts = ms.dof_points()
    ms.close()
    ms = MicroSolver(capi.BIO if bio else capi.ELLIPTIC, data, nc)
    ms.set_grid_locations(pts)
    ms.setup()
    N = len(pts)
    u, w = 1.0 + 0.3 * pts[:, 0], 0.5 - 0.2 * pts[:, 1]
    ms.set_macro_solution(u, w if bio else None)
    ms.run()
    fields = 2 if bio else 1
    ms.macro_setup(fields, kind, nc)
    systems = []
    for f in range(fields):
        A, C, b0, x0, rs, cl, av, cv = _macro_problem(pts, nc, dirichlet_all=not bio, rng=rng)
        ms.macro_set_system(f, f, rs, cl, av, cv, b0, x0)
        systems.append((A, C, b0))
    c = ms.coupling()
    c = list(c) if bio else [c]
    its = ms.macro_solve(1e-12, 10000, adopt=True)
    assert (its > 0).all() or nc == 1                # nc = 1: every DoF is a Dirichlet DoF, nothing to iterate
    want = [np.linalg.solve(A, b0 + C @ c[f]) for f, (A, C, b0) in enumerate(systems)]
    for f in range(fields):
        got = ms.macro_get_solution(f)
        assert np.abs(got - want[f]).max() <= 1e-10 * np.abs(want[f]).max()

    # adopt: the next micro run uses the new macro values without any upload
    ms.assemble_system()
    ms.solve()
    x_dev = ms.solutions()
    ms.close()

    ref = MicroSolver(capi.BIO if bio else capi.ELLIPTIC, data, nc)
    ref.set_grid_locations(pts)
    ref.setup()
    ref.set_macro_solution(u, w if bio else None)
    ref.run()                                       # same history as above (bio warm-starts from its iterate)
    ref.set_macro_solution(want[0], want[1] if bio else None)
    ref.run()
    x_ref = ref.solutions()
    ref.close()
    assert np.abs(x_dev - x_ref).max() <= 1e-9 * np.abs(x_ref).max()


def test_host_given_right_hand_side_and_warm_start():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    rng = np.random.default_rng(3)
    nc = 5
    data = BioData(os.path.join(INPUT, "biocase.prm"))
    ms = MicroSolver(capi.BIO, data, nc)
    ms.set_grid_locations(np.zeros((1, 2)))
    ms.setup()
    pts = ms.dof_points()
    ms.close()
    ms = MicroSolver(capi.BIO, data, 2)
    ms.set_grid_locations(pts)
    ms.setup()
    ms.macro_setup(1, capi.MESH_SUBDIVIDED, nc)
    A, C, b0, x0, rs, cl, av, cv = _macro_problem(pts, nc, dirichlet_all=True, rng=rng)
    ms.macro_set_system(0, 0, rs, cl, av, None, b0, x0)
    b = rng.standard_normal(len(pts))
    ms.macro_set_rhs(0, b)
    its1 = ms.macro_solve(1e-12, 1000, adopt=False)
    x = ms.macro_get_solution(0)
    assert np.abs(x - np.linalg.solve(A, b)).max() < 1e-10
    ms.macro_set_rhs(0, b)
    its2 = ms.macro_solve(1e-12, 1000, adopt=False)       # warm start from the solution: converged at once
    assert its2[0] <= 1 < its1[0]
    # a matrix that is not symmetric is refused
    av_bad = av.copy()
    av_bad[1] += 1.0
    with pytest.raises(capi.MsgpuError):
        ms.macro_set_system(0, 0, rs, cl, av_bad, None, b0, x0)
    ms.close()
