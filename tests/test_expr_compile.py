"""The product's expression compiler + device VM (run on the CPU through tests/native) against the
oracle's independent tree-walking evaluator and against Python's own arithmetic."""
import math

import numpy as np
import pytest

from helpers import (ELLIPTIC, FN_BC, FN_JACOBIAN, FN_MAPPING, FN_RHS, FN_SOLUTION, MESH_REFINE_GLOBAL, Emul)
from test_reference_known_answers import oracle_eval, product_eval

EXPRS = [
    "x0*y1 - x1^2 - 2*y0*y0",
    "-x1^2",                       # sign binds weaker than ^
    "2^3^2",                       # right associative
    "-2^2 + (-2)^2",
    "(x0 + 1.5)*(0.847*(0.909090909090909*y0 + 1)^2 - 0.363*(0.909090909090909*y1 + 1)^2)",
    "sqrt(0.16*abs(x1)^2 + abs(x1^2/2 - 1)^2)",
    "y0*sin(y1)* exp(x0^2 + x1^2)",
    "cos (0.785)*y0 - sin( 0.785)*y1",   # blanks after function names
    "min(y0, y1, x0) + max(1, 4*exp(-2*(x0*y1))*sin(1)^2)",
    "if(y0 > 0.25, y1, -y1) + (x0 < x1) + (y0 >= y1) * 3",
    "sum(x0, x1, y0, y1) / avg(1, 2, 3) + sign(y0 - 0.5) + rint(3.5*y1) + int(2.6) + floor(-1.5) + ceil(1.2)",
    "1/(1+y0^2) + y1^4 + y0^3 - pow(2 + y1, 1.5) + atan2(y0, 2 + y1)",
    "(y0 > 0) & (y1 > 0) | (x0 > 10)",
    "log(2 + y0) + ln(2 + y1) + log2(3 + y0) + log10(3 + y1) + tanh(y0) + sinh(y1) + cosh(y0*y1)",
    "cot(1 + 0.2*y0) + sec(y1/2) + csc(1.3 + 0.1*y0) + erfc(y1) + asin(y0/2) + acos(y1/2) + atan(y0)",
    "1.6*2*(y0+2)",
    "D_2*(0.66896473162245*y0*y1 - 0.66896473162245*y0*(1 - y1)) - kappa_1*x1*sin(x0) + kappa_2*(x0 + x1 + y0*y1*(1 - y1))",
    "_pi * y0 + _e",
    "3e-1*y0 + .5*y1 + 2.*x0",
]
CONSTS = {"D_2": 1.25, "kappa_1": 0.7, "kappa_2": 1.3}
POINTS = [((0.2, 0.4), (0.0, 1.0)), ((-0.75, 0.5), (0.3, -0.9)), ((1.0, -1.0), (0.861136, 0.339981))]


@pytest.mark.parametrize("expr", EXPRS)
def test_product_vm_matches_oracle_evaluator(expr):
    for x, y in POINTS:
        ref = oracle_eval(expr, [*x, *y], CONSTS)
        got = product_eval(expr, x, y, CONSTS)
        assert got == pytest.approx(ref, rel=1e-14, abs=1e-15), (expr, x, y)


def test_python_cross_check():
    x, y = (0.2, 0.4), (0.3, -0.9)
    v = product_eval("-x1^2 + 2^3^2 + y0*sin(y1)*exp(x0^2 + x1^2)", x, y)
    assert v == pytest.approx(-(0.4 ** 2) + 2 ** 9 + 0.3 * math.sin(-0.9) * math.exp(0.2 ** 2 + 0.4 ** 2), rel=1e-14)


@pytest.mark.parametrize("bad", ["x0 +", "Abs(x0)", "zoo*(y0)", "x0 + t", "rand()", "(x0", "x0 $ x1", "sin(x0, x1)"])
def test_errors_are_reported_by_both(bad):
    with pytest.raises(ValueError):
        oracle_eval(bad, [0, 0, 0, 0])
    with pytest.raises(RuntimeError):
        product_eval(bad, (0, 0), (0, 0))


def test_line_tables_and_point_program_agree_with_direct_evaluation():
    """The split into system / line / point programs must not change values: the volume program's
    outputs through the tables equal the oracle's evaluation at the same quadrature point."""
    e = Emul(ELLIPTIC, MESH_REFINE_GLOBAL, 4, 12)
    mapping = "(x0 + 1.5)*(0.847*(0.909090909090909*y0 + 1)^2 - 0.363*(0.909090909090909*y1 + 1)^2);-0.242*(0.909090909090909*y0 + 1)^2 + 1.573*(0.909090909090909*y1 + 1)^2"
    jac = "(x0 + 1.5)*(1.4*y0 + 1.54);(x0 + 1.5)*(-0.6*y1 - 0.66);-0.4*y0 - 0.44;2.6*y1 + 2.86"
    rhs = "y0*sin(y1)*exp(x0^2 + x1^2) - sin(x0*x1) - cos(x0 + x1)"
    e.set_function(FN_MAPPING, mapping)
    e.set_function(FN_JACOBIAN, jac)
    e.set_function(FN_RHS, rhs)
    e.set_function(FN_BC, "y0")
    e.set_function(FN_SOLUTION, "y0")
    xy = np.array([[0.25, -0.5], [-1.0, 1.0]])
    e.setup(xy)
    from helpers import dptr, load_oracle
    lib = load_oracle()
    gx, gw = np.zeros(12), np.zeros(12)
    lib.orc_gauss01(12, dptr(gx), dptr(gw))
    h = 0.5
    n_slots, n0, n1, _ = e.stats()
    assert n0 >= 2 and n1 >= 2            # the Jacobian of this map is separable
    for k, x in enumerate(xy):
        for (cx, cy, qx, qy) in [(0, 0, 0, 0), (3, 1, 11, 4), (2, 3, 7, 7)]:
            y = (-1 + cx * h + h * gx[qx], -1 + cy * h + h * gx[qy])
            out = e.volume_point(k, cx, cy, qx, qy, 5)
            F = [oracle_eval(c, [*x, *y]) for c in jac.split(";")]
            zeta = [oracle_eval(c, [*x, *y]) for c in mapping.split(";")]
            g = oracle_eval(rhs, [*x, *zeta])
            np.testing.assert_allclose(out[:4], F, rtol=1e-15, atol=0)
            assert out[4] == pytest.approx(g, rel=1e-14)
    e.close()


def test_constant_jacobian_needs_no_tables():
    e = Emul(ELLIPTIC, MESH_REFINE_GLOBAL, 2, 12)
    e.set_function(FN_MAPPING, "1.5*y0 + 0.7*y1;-0.2*y0 + 2.3*y1")
    e.set_function(FN_JACOBIAN, "1.50000000000000;0.700000000000000;-0.200000000000000;2.30000000000000")
    e.set_function(FN_RHS, "5*y0*sin(y1) - 5")
    e.set_function(FN_BC, "5*y0*sin(y1)")
    e.set_function(FN_SOLUTION, "5*y0*sin(y1)")
    e.setup(np.array([[0.0, 0.0]]))
    n_slots, n0, n1, vol_len = e.stats()
    assert n_slots == 0
    text = e.disassemble(3)
    assert "LD c[" in text and "ST R[" in text
    e.close()
