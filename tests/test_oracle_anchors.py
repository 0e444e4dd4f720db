"""The oracle against the only independent numbers available for the deal.II part of the path:
the anchors of SURVEY.md §9, obtained there with a separate NumPy/SciPy restatement using sparse
DIRECT solves (nothing comes from a deal.II run: "parity unpinned").  Equal cycle
counts and 3-4 significant digits are what two correct restatements must share."""
import pytest

from helpers import OracleBio, OracleElliptic, OracleParabolic


@pytest.mark.parametrize("prm,refs,rows", [
    ("elliptic_linear.prm", (1, 2), [(5.059e-02, 1.822e-02), (4.969e-02, 1.816e-02), (4.969e-02, 1.816e-02)]),
    ("elliptic_linear.prm", (2, 3), [(3.951e-02, 4.590e-03), (1.346e-02, 4.477e-03), (1.346e-02, 4.476e-03)]),
])
def test_elliptic_full_history(prm, refs, rows):
    o = OracleElliptic(prm, *refs)
    assert o.run() == len(rows)          # stops after 3 cycles at every level
    for c, (ML2, mL2) in enumerate(rows):
        h = o.history(c, solutions=False)["errors"]
        assert h[2] == pytest.approx(ML2, rel=2e-3) and h[0] == pytest.approx(mL2, rel=2e-3)
    assert "cycle cells dofs" in o.table()
    o.close()


def test_elliptic_non_linear_and_rot_map():
    o = OracleElliptic("elliptic_non_linear.prm", 2, 3)
    assert o.run() == 3
    h = o.history(2, solutions=False)["errors"]
    assert h[2] == pytest.approx(1.362e-02, rel=2e-3) and h[0] == pytest.approx(2.290e-03, rel=2e-3)
    o.close()
    o = OracleElliptic("rot_map_test.prm", 1, 2)
    assert o.run() == 4                  # converges, stops at cycle index 3
    h = o.history(3, solutions=False)["errors"]
    assert h[2] == pytest.approx(2.172e-01, rel=2e-3) and h[0] == pytest.approx(1.797e-01, rel=2e-3)
    o.close()


def test_elliptic_default_input_diverges():
    """map_test.prm, the reference's default, blows up by a factor 400-600 per cycle."""
    o = OracleElliptic("map_test.prm", 1, 2)
    o.run(cycle_limit=3)
    res = [o.history(c, solutions=False)["errors"] for c in range(3)]
    growth = (res[2][0] + res[2][2]) / (res[1][0] + res[1][2])
    assert 100 < growth < 2000
    o.close()


def test_bio_linear_converges_in_21_cycles():
    o = OracleBio("linear.prm", 4, 16, threads=8)
    assert o.run() == 21
    h = o.history(20, solutions=False)
    assert h["macro_l2"] == pytest.approx(3.455e-02, rel=2e-3) and h["micro_l2"] == pytest.approx(2.905e-02, rel=2e-3)
    o.close()


@pytest.mark.parametrize("prm,passes,rows", [
    ("linear_time.prm", 19, {0: (5.425e-02, 1.464e-02), 1: (3.664e-02, 1.270e-02)}),
    ("full_nonlinear.prm", 6, {0: (4.103e-03, 2.739e-01), 1: (2.136e-03, 1.483e-01)}),
])
def test_parabolic_sweep_levels(prm, passes, rows):
    for level, (ML2, mL2) in rows.items():
        h_inv = round(8 * 2 ** (level / 2.0))
        t_inv = round(4 * 2 ** level)
        o = OracleParabolic(prm, h_inv, h_inv, t_inv)
        steps = o.run()
        assert steps == t_inv // 2
        assert o.first_step_passes() == passes
        last = o.history(steps - 1, solutions=False)
        assert last["macro_l2"] == pytest.approx(ML2, rel=2e-3) and last["micro_l2"] == pytest.approx(mL2, rel=2e-3)
        o.close()
