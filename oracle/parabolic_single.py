"""TEST INFRASTRUCTURE — never imported by the product (dealiiscale_b200/).

NumPy/SciPy restatement of ONE micro grid of the elliptic-parabolic driver's `RhoSolver`
(`src/elliptic-parabolic/rho_solver.cpp:55-93` set-up and L2 projection of `init_rho`, `:108-218` implicit-Euler
assembly with Robin faces left / right and Neumann faces top / bottom, `:222-229` solve) and of `PiSolver::get_micro_mass`
(`src/elliptic-parabolic/pi_solver.cpp:178-197`).  Independent of `oracle/*.h` (own expression evaluation, quadrature,
sparse direct solves).  Third "second opinion" next to robin_single.py and dirichlet_single.py.

Parity: unpinned by reference golden values; pinned analytically by `input/linear_time.prm`, whose micro solution
`3 y0 + y1 + 2` is in the Q1 space and stationary: with the exact macro value the discrete solution must reproduce its
nodal values to rounding at every time step.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from robin_single import _shape, gauss01

_NS = {k: getattr(np, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "sinh", "cosh", "tanh")}


def evaluate(expr, consts, x, y, t):
    v = eval(expr.replace("^", "**"), {"__builtins__": {}},  # noqa: S307 (test input files)
             dict(_NS, **consts, x0=x[0], x1=x[1], y0=y[0], y1=y[1], t=t, abs=np.abs))
    return np.broadcast_to(np.asarray(v, dtype=np.float64), np.shape(y[0])).copy()


class RhoGrid:
    def __init__(self, data, consts, x, h_inv):
        """data: dict with micro_rhs, micro_bc_robin, micro_bc_neumann, init_rho; consts: A, D, theta, kappa, p_F, R."""
        self.d, self.c, self.x = data, consts, x
        self.nc, self.m = h_inv, h_inv + 1
        self.n = self.m * self.m
        self.h = 2.0 / h_inv
        self._matrices()
        self.rho = self.project(self.d["init_rho"], 0.0, 3)   # rho_solver.cpp:80-83 (init_rho has no time variable)

    def node(self, ix, iy):
        return iy * self.m + ix

    def points(self):
        g = -1.0 + self.h * np.arange(self.m)
        X, Y = np.meshgrid(g, g, indexing="xy")
        return np.stack([X.ravel(), Y.ravel()], axis=1)

    def _cells(self):
        for cy in range(self.nc):
            for cx in range(self.nc):
                yield cx, cy, [self.node(cx, cy), self.node(cx + 1, cy), self.node(cx, cy + 1), self.node(cx + 1, cy + 1)]

    def _matrices(self):
        h = self.h
        xi, w = gauss01(2)                                    # MatrixCreator with QGauss(2) (rho_solver.cpp:64-65)
        XI, ETA = np.meshgrid(xi, xi, indexing="ij")
        W = np.outer(w, w) * h * h
        v, gx, gy = _shape(XI, ETA)
        Ml = np.array([[np.sum(v[i] * v[j] * W) for j in range(4)] for i in range(4)])
        Ll = np.array([[np.sum((gx[i] * gx[j] + gy[i] * gy[j]) / (h * h) * W) for j in range(4)] for i in range(4)])
        M = np.zeros((self.n, self.n))
        L = np.zeros((self.n, self.n))
        for _, _, dofs in self._cells():
            M[np.ix_(dofs, dofs)] += Ml
            L[np.ix_(dofs, dofs)] += Ll
        self.M, self.L = M, L

    def load(self, expr, t, q):
        """VectorTools::create_right_hand_side / the load of VectorTools::project with QGauss(q)."""
        h = self.h
        xi, w = gauss01(q)
        XI, ETA = np.meshgrid(xi, xi, indexing="ij")
        W = np.outer(w, w) * h * h
        v, _, _ = _shape(XI, ETA)
        b = np.zeros(self.n)
        for cx, cy, dofs in self._cells():
            f = evaluate(expr, self.c, self.x, (-1.0 + cx * h + h * XI, -1.0 + cy * h + h * ETA), t)
            for i in range(4):
                b[dofs[i]] += np.sum(f * v[i] * W)
        return b

    def project(self, expr, t, q):
        return np.linalg.solve(self.M, self.load(expr, t, q))

    def step(self, pi_k, dt, t):
        """One implicit-Euler step to time t (euler = 1, rho_solver.cpp:127-216)."""
        kappa, p_F, R, D = self.c["kappa"], self.c["p_F"], self.c["R"], self.c["D"]
        h = self.h
        b = self.M @ self.rho + dt * self.load(self.d["micro_rhs"], t, 2)
        A = self.M + dt * D * self.L
        s, w = gauss01(2)
        JxW = w * h
        for cx, cy, dofs in self._cells():
            faces = []
            if cx == 0:
                faces.append((True, np.zeros_like(s), s))           # x-min: Robin
            if cx == self.nc - 1:
                faces.append((True, np.ones_like(s), s))            # x-max: Robin
            if cy == 0:
                faces.append((False, s, np.zeros_like(s)))          # |y1| = 1: Neumann (make_grid, :40-50)
            if cy == self.nc - 1:
                faces.append((False, s, np.ones_like(s)))
            for robin, fxi, feta in faces:
                vf, _, _ = _shape(fxi, feta)
                Y = (-1.0 + cx * h + h * fxi, -1.0 + cy * h + h * feta)
                if robin:
                    for i in range(4):
                        for j in range(4):
                            A[dofs[i], dofs[j]] += kappa * dt * R * np.sum(vf[i] * vf[j] * JxW)
                    g = kappa * (pi_k + p_F) + evaluate(self.d["micro_bc_robin"], self.c, self.x, Y, t)
                else:
                    g = evaluate(self.d["micro_bc_neumann"], self.c, self.x, Y, t)
                for i in range(4):
                    b[dofs[i]] += np.sum(g * dt * vf[i] * JxW)
        self.A, self.b = A, b
        self.rho = spla.spsolve(sp.csc_matrix(A), b)
        return self.rho

    def mass(self):
        """PiSolver::get_micro_mass: int rho dy with QGauss(2) (exact for Q1)."""
        return float(np.ones(self.n) @ (self.M @ self.rho))
