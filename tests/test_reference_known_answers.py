"""The reference's own unit tests (tests/test_tools.cpp), run against the oracle AND against the
product's expression compiler + VM (through the CPU harness).  These are the only golden values the
reference holds for the micro path: two parser values and five MapMap behaviours."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import ELLIPTIC, FN_BC, FN_JACOBIAN, FN_MAPPING, FN_RHS, FN_SOLUTION, MESH_REFINE_GLOBAL, Emul, dptr, load_oracle

PI = {"pi": math.pi}


def oracle_eval(expr, values, constants=None, variables="x0,x1,y0,y1"):
    lib = load_oracle()
    constants = constants or {}
    names = [k.encode() for k in constants]
    arr = (C.c_char_p * max(1, len(names)))(*names)
    vals = np.array(list(constants.values()) or [0.0])
    v = np.array(values, dtype=np.float64)
    out = C.c_double()
    rc = lib.orc_eval(expr.encode(), variables.encode(), arr, dptr(vals), len(names), dptr(v), C.byref(out))
    if rc != 0:
        raise ValueError(lib.orc_last_error().decode())
    return out.value


def product_eval(expr, x, y, constants=None):
    """Value of `expr` at macro point x, micro point y through the product's compiler and VM
    (identity mapping, expression in the exact-solution slot, arbitrary-point program)."""
    e = Emul(ELLIPTIC, MESH_REFINE_GLOBAL, 2, 2)
    e.set_function(FN_MAPPING, "y0;y1")
    e.set_function(FN_JACOBIAN, "1;0;0;1")
    e.set_function(FN_RHS, "0")
    e.set_function(FN_BC, "0")
    e.set_function(FN_SOLUTION, expr, constants or {})
    e.setup(np.array([x], dtype=np.float64))
    v = e.free_point(0, 0, y[0], y[1])
    e.close()
    return v


# ---- tests/test_tools.cpp:21-29 ------------------------------------------------------------------
def test_function_mvalue():
    expr, x, y = "x0*y1 - x1^2 - 2*y0*y0", (0.2, 0.4), (0.0, 1.0)
    expected = 0.2 - 0.16
    assert oracle_eval(expr, [*x, *y], PI) == pytest.approx(expected, rel=1e-11)
    assert product_eval(expr, x, y, PI) == pytest.approx(expected, rel=1e-11)


# ---- tests/test_tools.cpp:31-40 ------------------------------------------------------------------
def test_function_set_x():
    expr, x, y = "cos(pi*x0*y1) - sin(x1*2*pi) - 2", (0.5, -2.0), (-3.0, 1.0)
    assert oracle_eval(expr, [*x, *y], PI) == pytest.approx(-2.0, rel=1e-11)
    assert product_eval(expr, x, y, PI) == pytest.approx(-2.0, rel=1e-11)


# ---- tests/test_tools.cpp:45-103 (MapMap) -----------------------------------------------------------
class _MapMap:
    def __init__(self):
        self.lib = load_oracle()
        self.lib.orc_mapmap_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.orc_mapmap_create())

    def set(self, px, py, det, kkt):
        px, py, kkt = (np.array(a, dtype=np.float64) for a in (px, py, kkt))
        self.lib.orc_mapmap_set(self.h, dptr(px), len(px), dptr(py), len(py), C.c_double(det), dptr(kkt))

    def get(self, px, py):
        px, py = (np.array(a, dtype=np.float64) for a in (px, py))
        det, kkt = C.c_double(), np.zeros(3)
        if self.lib.orc_mapmap_get(self.h, dptr(px), len(px), dptr(py), len(py), C.byref(det), dptr(kkt)) != 0:
            raise KeyError(self.lib.orc_last_error().decode())
        return det.value, kkt

    def size(self):
        return self.lib.orc_mapmap_size(self.h)


PX, PY, KKT = (2.1415, 2.71812, 4.532), (3.1415, 2.71812), (2.0, 1.0, 2.0)


def test_mapmap_setting():
    m = _MapMap()
    m.set(PX, PY, 3, KKT)
    assert m.size() == 1


def test_mapmap_duplicate_setting():
    m = _MapMap()
    m.set(PX, PY, 3, KKT)
    m.set(PX, PY, 3, KKT)
    assert m.size() == 1


def test_mapmap_getting():
    m = _MapMap()
    m.set(PX, PY, 3, KKT)
    det, kkt = m.get(PX, PY)
    assert kkt[0] == 2 and det == 3


def test_mapmap_invalid_getting():
    m = _MapMap()
    with pytest.raises(KeyError):
        m.get(PX, PY)


def test_mapmap_string_robustness():
    m = _MapMap()
    m.set(PX, PY, 3, KKT)
    m.set((2.14150, 2.7181200, 4.53200000000), PY, 3, KKT)
    assert m.size() == 1


def test_mapmap_key_is_collision_free_on_the_quadrature_lattice():
    """at 6 significant digits no two distinct (macro DoF, quadrature point)
    pairs share a key, so the cache is semantically the dense table the product recomputes."""
    lib = load_oracle()
    x = np.zeros(12)
    w = np.zeros(12)
    lib.orc_gauss01(12, dptr(x), dptr(w))
    nc = 32
    h = 2.0 / nc
    pts = sorted({float(f"{-1.0 + c * h + h * q:.6g}") for c in range(nc) for q in x})
    assert len(pts) == nc * 12
