"""TEST INFRASTRUCTURE — never imported by the product (dealiiscale_b200/).

NumPy/SciPy restatement of the single pulled-back Robin system of the reference's playground program
`src/playground/transform.cpp` (default input `input/parsed_mapping.prm`, `transform.cpp:491`).  One such
system is exactly one micro system of the bio driver (`src/bio-multiscale/micro.cpp:183-276`) with
kappa_2 = kappa_4 = D_2 = 1 and kappa_1 = kappa_3 = 0, which makes it an independent cross-check of the
Robin / Neumann assembly with a y-dependent Jacobian.  Independent of `oracle/*.h`: its own
expression evaluation (Python `eval` on NumPy arrays), its own quadrature and a sparse direct solve.

Parity: unpinned by reference golden values (the reference holds none for this program); pinned
analytically — the discrete solution must approach `ref_solution` of the .prm with order 2
(`transform.cpp:377-386` measures exactly that error).
"""
from __future__ import annotations

import re

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# transform.cpp:102-116: declared entries and their defaults
DEFAULTS = {
    "solution": "sin(x)*y",
    "rhs": "sin(x)*y",
    "ref_solution": "(x/2 + sqrt(3)*y/2)*sin(sqrt(3)*x/2 - y/2)",
    "jac_mapping": "cos(pi/6);-sin(pi/6);sin(pi/6);cos(pi/6)",
    "mapping": "x*cos(pi/6) - y*sin(pi/6);x*sin(pi/6) + y*cos(pi/6)",
    "left_robin": "y*sin(x) - sqrt(3)*y*cos(x)/2 - sin(x)/2",
    "right_robin": "y*sin(x) + sqrt(3)*y*cos(x)/2 + sin(x)/2",
    "up_neumann": "-y*cos(x)/2 + sqrt(3)*sin(x)/2",
    "down_neumann": "y*cos(x)/2 - sqrt(3)*sin(x)/2",
}


def read_prm(path):
    """ParameterHandler subset: `set key = value`, `#` comments, undeclared key = error (transform.cpp:117)."""
    out = dict(DEFAULTS)
    for line in open(path):
        line = line.split("#", 1)[0].strip()
        if not line:
            continue
        mt = re.match(r"set\s+(\w+)\s*=\s*(.*)$", line)
        if not mt or mt.group(1) not in out:
            raise ValueError(f"{path}: cannot parse '{line}'")
        out[mt.group(1)] = mt.group(2).strip()
    return out


_NS = {k: getattr(np, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "sinh", "cosh", "tanh")}
_NS["pi"] = np.pi  # transform.cpp:119-120


def components(expr):
    """';'-separated components (a trailing ';' is tolerated, as FunctionParser does for `jac_mapping = 2;0;0;3;`)."""
    parts = [p.strip() for p in expr.split(";")]
    return [p for p in parts if p]


def evaluate(expr, x, y):
    """muParser expression in the variables x, y -> list of arrays (one per component)."""
    res = []
    for c in components(expr):
        v = eval(c.replace("^", "**"), {"__builtins__": {}}, dict(_NS, x=x, y=y))  # noqa: S307 (test input files)
        res.append(np.broadcast_to(np.asarray(v, dtype=np.float64), np.shape(x)).copy())
    return res


def gauss01(p):
    """QGauss<1>(p) on [0, 1]."""
    t, w = np.polynomial.legendre.leggauss(p)
    return 0.5 * (t + 1.0), 0.5 * w


def _shape(xi, eta):
    """FE_Q<2>(1) on the unit cell, vertex order (0,0),(1,0),(0,1),(1,1): values [4,...] and reference gradients."""
    v = np.array([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta])
    gx = np.array([-(1 - eta), (1 - eta), -eta, eta])
    gy = np.array([-(1 - xi), -xi, (1 - xi), xi])
    return v, gx, gy


class RobinSystem:
    """`RobinSolver` of transform.cpp on hyper_cube(-1,1) with `nc` cells per axis.  Nodes are numbered
    row-major, x fastest (the numbering is irrelevant to the comparison: tests match nodes by coordinates)."""

    def __init__(self, prm, nc, q_volume=8, q_face=8):
        self.p = prm if isinstance(prm, dict) else read_prm(prm)
        self.nc, self.m = nc, nc + 1
        self.n = self.m * self.m
        self.h = 2.0 / nc
        self.qv, self.qf = q_volume, q_face
        self.A = None
        self.b = None
        self.x = None

    def node(self, ix, iy):
        return iy * self.m + ix

    def points(self):
        g = -1.0 + self.h * np.arange(self.m)
        X, Y = np.meshgrid(g, g, indexing="xy")
        return np.stack([X.ravel(), Y.ravel()], axis=1)

    def assemble(self):
        """transform.cpp:263-361."""
        nc, h = self.nc, self.h
        xi, w = gauss01(self.qv)
        XI, ETA = np.meshgrid(xi, xi, indexing="ij")
        W = np.outer(w, w) * h * h  # JxW
        v, gx, gy = _shape(XI, ETA)
        gx, gy = gx / h, gy / h
        rows, cols, vals = [], [], []
        b = np.zeros(self.n)
        sf, wf = gauss01(self.qf)
        for cy in range(nc):
            for cx in range(nc):
                x0, y0 = -1.0 + cx * h, -1.0 + cy * h
                dofs = [self.node(cx, cy), self.node(cx + 1, cy), self.node(cx, cy + 1), self.node(cx + 1, cy + 1)]
                X, Y = x0 + h * XI, y0 + h * ETA
                F = evaluate(self.p["jac_mapping"], X, Y)  # row-major F[i][j] = comp[2 i + j]   (transform.cpp:52-60)
                det = F[0] * F[3] - F[1] * F[2]
                # K = F^-1 F^-T                                                              (transform.cpp:154-159)
                i00, i01, i10, i11 = F[3] / det, -F[1] / det, -F[2] / det, F[0] / det
                k00 = i00 * i00 + i01 * i01
                k01 = i00 * i10 + i01 * i11
                k11 = i10 * i10 + i11 * i11
                zx, zy = evaluate(self.p["mapping"], X, Y)
                f = evaluate(self.p["rhs"], zx, zy)[0]  # rhs(map(q))                            (transform.cpp:295)
                Aloc = np.zeros((4, 4))
                bloc = np.zeros(4)
                for i in range(4):
                    for j in range(4):
                        Aloc[i, j] = np.sum(
                            (gx[i] * (k00 * gx[j] + k01 * gy[j]) + gy[i] * (k01 * gx[j] + k11 * gy[j])) * W * det)
                    bloc[i] = np.sum(v[i] * f * det * W)
                # boundary faces: 0 = x-min (LEFT_ROBIN), 1 = x-max (RIGHT_ROBIN), 2 = y-min (DOWN), 3 = y-max (UP)
                # (ids are set from the face centres, transform.cpp:226-243; the bio driver uses the same four ids)
                faces = []
                if cx == 0:
                    faces.append(("left_robin", True, np.zeros_like(sf), sf, (-1.0, 0.0)))
                if cx == nc - 1:
                    faces.append(("right_robin", True, np.ones_like(sf), sf, (1.0, 0.0)))
                if cy == 0:
                    faces.append(("down_neumann", False, sf, np.zeros_like(sf), (0.0, -1.0)))
                if cy == nc - 1:
                    faces.append(("up_neumann", False, sf, np.ones_like(sf), (0.0, 1.0)))
                for name, robin, fxi, feta, nrm in faces:
                    Xf, Yf = x0 + h * fxi, y0 + h * feta
                    Ff = evaluate(self.p["jac_mapping"], Xf, Yf)
                    # |F R n|, R = [[0,-1],[1,0]]                                            (transform.cpp:305-309)
                    rn = (-nrm[1], nrm[0])
                    t0 = Ff[0] * rn[0] + Ff[1] * rn[1]
                    t1 = Ff[2] * rn[0] + Ff[3] * rn[1]
                    jb = np.sqrt(t0 * t0 + t1 * t1)
                    zfx, zfy = evaluate(self.p["mapping"], Xf, Yf)
                    g = evaluate(self.p[name], zfx, zfy)[0]
                    vf, _, _ = _shape(fxi, feta)
                    JxW = wf * h
                    for i in range(4):
                        if robin:
                            for j in range(4):
                                Aloc[i, j] += np.sum(vf[i] * vf[j] * jb * JxW)
                        bloc[i] += np.sum(g * jb * vf[i] * JxW)
                for i in range(4):
                    for j in range(4):
                        rows.append(dofs[i])
                        cols.append(dofs[j])
                        vals.append(Aloc[i, j])
                    b[dofs[i]] += bloc[i]
        self.A = sp.csr_matrix((vals, (rows, cols)), shape=(self.n, self.n))
        self.b = b
        return self

    def solve(self):
        """transform.cpp:363-367 runs CG to 1e-10; a sparse direct solve is the converged limit of that."""
        self.x = spla.spsolve(self.A.tocsc(), self.b)
        return self.x

    def flux(self, side, kappa=1.0, x=None):
        """The micro -> macro flux integral of the bio driver over the inflow (side 0, x-min) or outflow (side 1, x-max)
        boundary: sum over faces and points of (-kappa v_h + bc(zeta)) JxW |F R n|  (src/bio-multiscale/macro.cpp:254-296);
        bc is this system's left / right Robin function."""
        x = self.x if x is None else x
        nc, h = self.nc, self.h
        sf, wf = gauss01(self.qf)
        cx = 0 if side == 0 else nc - 1
        fxi = np.zeros_like(sf) if side == 0 else np.ones_like(sf)
        nrm = (-1.0, 0.0) if side == 0 else (1.0, 0.0)
        name = "left_robin" if side == 0 else "right_robin"
        tot = 0.0
        for cy in range(nc):
            x0, y0 = -1.0 + cx * h, -1.0 + cy * h
            dofs = [self.node(cx, cy), self.node(cx + 1, cy), self.node(cx, cy + 1), self.node(cx + 1, cy + 1)]
            Xf, Yf = x0 + h * fxi, y0 + h * sf
            Ff = evaluate(self.p["jac_mapping"], Xf, Yf)
            rn = (-nrm[1], nrm[0])
            jb = np.sqrt((Ff[0] * rn[0] + Ff[1] * rn[1]) ** 2 + (Ff[2] * rn[0] + Ff[3] * rn[1]) ** 2)
            zfx, zfy = evaluate(self.p["mapping"], Xf, Yf)
            g = evaluate(self.p[name], zfx, zfy)[0]
            vf, _, _ = _shape(fxi, sf)
            vh = sum(x[dofs[i]] * vf[i] for i in range(4))
            tot += np.sum((-kappa * vh + g) * wf * h * jb)
        return float(tot)

    def l2_error(self, x=None, q=5):
        """integrate_difference against ref_solution with QGauss(5), L2 norm (transform.cpp:377-386)."""
        x = self.x if x is None else x
        nc, h = self.nc, self.h
        xi, w = gauss01(q)
        XI, ETA = np.meshgrid(xi, xi, indexing="ij")
        W = np.outer(w, w) * h * h
        v, _, _ = _shape(XI, ETA)
        tot = 0.0
        for cy in range(nc):
            for cx in range(nc):
                x0, y0 = -1.0 + cx * h, -1.0 + cy * h
                dofs = [self.node(cx, cy), self.node(cx + 1, cy), self.node(cx, cy + 1), self.node(cx + 1, cy + 1)]
                uh = sum(x[dofs[i]] * v[i] for i in range(4))
                ex = evaluate(self.p["ref_solution"], x0 + h * XI, y0 + h * ETA)[0]
                tot += np.sum((uh - ex) ** 2 * W)
        return np.sqrt(tot)
