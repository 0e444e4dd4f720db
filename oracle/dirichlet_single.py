"""TEST INFRASTRUCTURE — never imported by the product (dealiiscale_b200/).

NumPy/SciPy restatement of ONE micro system of the elliptic-elliptic driver
(`src/elliptic-elliptic/micro.cpp:108-171`: pulled-back stiffness, load `(u_k + g(x_k, zeta(x_k, y))) phi_i J`, Dirichlet
values `micro_bc(x_k, zeta(x_k, y_b))`, `MatrixTools::apply_boundary_values`) and of its bulk integral
(`src/elliptic-elliptic/macro.cpp:134-157`).  Independent of `oracle/*.h`: expressions are evaluated by Python on NumPy
arrays, quadrature and elimination are written out again, the system is solved directly.  Second opinion next to the
C++ oracle ("cheap insurance against oracle bugs").

Parity: unpinned by reference golden values (there are none for this path); pinned analytically through the
manufactured micro solutions of the parameter files (`micro_solution`), which the Q1 solution must approach with
order 2 once `u_k` is the exact macro value.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from robin_single import _shape, components, gauss01

_NS = {k: getattr(np, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "sinh", "cosh", "tanh")}


def evaluate(expr, x, y):
    """Expression in x0, x1 (macro point) and y0, y1 (micro point, arrays) -> list of component arrays."""
    out = []
    for c in components(expr):
        v = eval(c.replace("^", "**"), {"__builtins__": {}}, dict(_NS, x0=x[0], x1=x[1], y0=y[0], y1=y[1], abs=np.abs))  # noqa: S307
        out.append(np.broadcast_to(np.asarray(v, dtype=np.float64), np.shape(y[0])).copy())
    return out


class DirichletSystem:
    """One grid of `MicroSolver` (elliptic-elliptic) on hyper_cube(-1,1) with nc cells per axis; nodes row-major,
    x fastest (tests match nodes by coordinates)."""

    def __init__(self, data, x, u_k, nc, q=12):
        self.d, self.x, self.u, self.nc, self.q = data, x, u_k, nc, q  # data: dict with the micro_* / mapping strings
        self.m = nc + 1
        self.n = self.m * self.m
        self.h = 2.0 / nc

    def node(self, ix, iy):
        return iy * self.m + ix

    def points(self):
        g = -1.0 + self.h * np.arange(self.m)
        X, Y = np.meshgrid(g, g, indexing="xy")
        return np.stack([X.ravel(), Y.ravel()], axis=1)

    def assemble(self):
        nc, h = self.nc, self.h
        xi, w = gauss01(self.q)
        XI, ETA = np.meshgrid(xi, xi, indexing="ij")
        W = np.outer(w, w) * h * h
        v, gx, gy = _shape(XI, ETA)
        gx, gy = gx / h, gy / h
        A = np.zeros((self.n, self.n))
        b = np.zeros(self.n)
        self.jw = np.zeros(self.n)  # int phi_i det F: the weights of the bulk integral
        for cy in range(nc):
            for cx in range(nc):
                Y = (-1.0 + cx * h + h * XI, -1.0 + cy * h + h * ETA)
                dofs = [self.node(cx, cy), self.node(cx + 1, cy), self.node(cx, cy + 1), self.node(cx + 1, cy + 1)]
                F = evaluate(self.d["jac_mapping"], self.x, Y)   # row-major (multiscale_function_parser.cpp:398-403)
                det = F[0] * F[3] - F[1] * F[2]
                i00, i01, i10, i11 = F[3] / det, -F[1] / det, -F[2] / det, F[0] / det
                k00, k01, k11 = i00 * i00 + i01 * i01, i00 * i10 + i01 * i11, i10 * i10 + i11 * i11
                Z = evaluate(self.d["mapping"], self.x, Y)
                g = evaluate(self.d["micro_rhs"], self.x, Z)[0]   # mapper composition (pde_data.cpp:32)
                for i in range(4):
                    for j in range(4):
                        A[dofs[i], dofs[j]] += np.sum(
                            (gx[i] * (k00 * gx[j] + k01 * gy[j]) + gy[i] * (k01 * gx[j] + k11 * gy[j])) * W * det)
                    b[dofs[i]] += np.sum((self.u + g) * v[i] * W * det)   # micro.cpp:126-135
                    self.jw[dofs[i]] += np.sum(v[i] * W * det)
        # interpolate_boundary_values + apply_boundary_values (micro.cpp:165-169)
        pts = self.points()
        self.boundary = [i for i in range(self.n) if abs(abs(pts[i, 0]) - 1) < 1e-12 or abs(abs(pts[i, 1]) - 1) < 1e-12]
        Zb = evaluate(self.d["mapping"], self.x, (pts[self.boundary, 0], pts[self.boundary, 1]))
        vals = evaluate(self.d["micro_bc"], self.x, Zb)[0]
        self.x0 = np.zeros(self.n)
        for dnode, val in zip(self.boundary, vals):
            add = A[dnode, dnode]
            col = A[:, dnode].copy()
            A[dnode, :] = 0.0
            A[:, dnode] = 0.0
            A[dnode, dnode] = add
            b -= col * val
            b[dnode] = val * add
            self.x0[dnode] = val
        self.A, self.b = A, b
        return self

    def solve(self):
        self.sol = spla.spsolve(sp.csc_matrix(self.A), self.b)
        return self.sol

    def bulk(self):
        """get_micro_bulk (macro.cpp:134-157): int v_k det F dy."""
        return float(self.jw @ self.sol)

    def l2_error(self, q=12):
        """per-grid error against micro_solution composed with the mapping (micro.cpp:193-202)."""
        nc, h = self.nc, self.h
        xi, w = gauss01(q)
        XI, ETA = np.meshgrid(xi, xi, indexing="ij")
        W = np.outer(w, w) * h * h
        v, _, _ = _shape(XI, ETA)
        tot = 0.0
        for cy in range(nc):
            for cx in range(nc):
                Y = (-1.0 + cx * h + h * XI, -1.0 + cy * h + h * ETA)
                dofs = [self.node(cx, cy), self.node(cx + 1, cy), self.node(cx, cy + 1), self.node(cx + 1, cy + 1)]
                uh = sum(self.sol[dofs[i]] * v[i] for i in range(4))
                Z = evaluate(self.d["mapping"], self.x, Y)
                ex = evaluate(self.d["micro_solution"], self.x, Z)[0]
                tot += np.sum((uh - ex) ** 2 * W)
        return np.sqrt(tot)
