"""Differential test of the expression engine on generated expressions: the product's compiler (hash-consed DAG,
constant folding, split into per-system / per-line / per-point programs, accumulator bytecode) + VM against the
oracle's independent tree-walking evaluator of the muParser grammar (src/tools/multiscale_function_parser.cpp:209-233,
:338-430).  Deterministic (derandomised hypothesis), CPU only."""
import math

import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from test_reference_known_answers import oracle_eval, product_eval

VARS = ["x0", "x1", "y0", "y1"]
CONSTS = {"D_2": 1.25, "kappa_1": 0.7}
NUMS = ["0.5", "2", "3", "1.25", "0.1", "7e-1", ".25", "2.", "1e1"]

leaf = st.one_of(st.sampled_from(VARS), st.sampled_from(NUMS), st.sampled_from(list(CONSTS)))


def _combine(children):
    a_b = st.tuples(children, children)
    return st.one_of(
        a_b.map(lambda t: f"({t[0]} + {t[1]})"),
        a_b.map(lambda t: f"({t[0]} - {t[1]})"),
        a_b.map(lambda t: f"{t[0]}*{t[1]}"),
        a_b.map(lambda t: f"{t[0]}/(2 + sin({t[1]}))"),          # denominator in [1, 3]
        children.map(lambda e: f"-{e}"),                           # sign binds weaker than ^
        children.map(lambda e: f"({e})^2"),
        children.map(lambda e: f"({e})^3"),
        a_b.map(lambda t: f"2^sin({t[0]})^({t[1]})^2"),            # ^ is right associative
        children.map(lambda e: f"sin({e})"),
        children.map(lambda e: f"cos ({e})"),                      # blank before the parenthesis is legal
        children.map(lambda e: f"tanh({e})"),
        children.map(lambda e: f"exp(sin({e}))"),
        children.map(lambda e: f"sqrt(abs({e}))"),
        children.map(lambda e: f"log(1 + abs({e}))"),
        children.map(lambda e: f"atan({e})"),
        a_b.map(lambda t: f"min({t[0]}, {t[1]})"),
        a_b.map(lambda t: f"max({t[0]}, {t[1]}, 0.5)"),
        st.tuples(children, children, children).map(lambda t: f"if({t[0]} > {t[1]}, {t[2]}, -{t[2]})"),
        a_b.map(lambda t: f"(({t[0]} < {t[1]}) | ({t[0]} >= 1))"),
        a_b.map(lambda t: f"pow(1 + abs({t[0]}), sin({t[1]}))"),
        a_b.map(lambda t: f"atan2({t[0]}, 2 + cos({t[1]}))"),
    )


EXPR = st.recursive(leaf, _combine, max_leaves=12)
POINT = st.tuples(*[st.floats(min_value=-1.0, max_value=1.0, allow_nan=False, width=64)] * 4)


@settings(max_examples=300, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.filter_too_much, HealthCheck.data_too_large])
@given(expr=EXPR, p=POINT)
def test_generated_expressions_agree(expr, p):
    x, y = (p[0], p[1]), (p[2], p[3])
    ref = oracle_eval(expr, [*x, *y], CONSTS)
    got = product_eval(expr, x, y, CONSTS)
    if math.isnan(ref) or math.isinf(ref):
        assert math.isnan(got) or got == ref
    else:
        assert got == pytest.approx(ref, rel=1e-13, abs=1e-300), expr
