"""Host logic: .prm reader with ParameterHandler semantics (reference src/tools/pde_data.cpp)."""
import os

import pytest

from helpers import INPUT
from dealiiscale_b200.prm import BioData, EllipticData, ParameterHandler, PrmError, TwoPressureData


def test_elliptic_defaults_and_override(tmp_path):
    p = tmp_path / "a.prm"
    p.write_text("# comment\nset micro_rhs = -5   # trailing comment\n\nset mapping = y0*3;y1*2\n")
    d = EllipticData(str(p))
    assert d.micro["rhs"] == "-5"
    assert d.micro["mapping"] == "y0*3;y1*2"
    assert d.micro["jac_mapping"].startswith("1.6*(3+x0+x1)")        # declared default (pde_data.cpp:19-20)
    assert d.macro["solution"] == "sin(x0*x1) + cos(x0 + x1)"


def test_undeclared_key_aborts(tmp_path):
    p = tmp_path / "b.prm"
    p.write_text("set solution_u = 0\n")
    with pytest.raises(PrmError):
        EllipticData(str(p))                                           # bio key in an elliptic file
    with pytest.raises(PrmError):
        EllipticData(os.path.join(INPUT, "linear.prm"))               # the committed linear.prm is a bio set
    with pytest.raises(PrmError):
        BioData(os.path.join(INPUT, "map_test.prm"))


def test_bio_constants_and_silent_defaults():
    d = BioData(os.path.join(INPUT, "biocase.prm"))
    assert d.constants["D_1"] == 0.1 and d.constants["kappa_1"] == 1.0 and d.constants["D_2"] == 1.0
    assert not d.has_solution
    # biocase.prm sets neither bulk_rhs_v nor solution_v: the declared defaults apply (pde_data.cpp:55-56)
    assert d.micro["rhs"] == "y0*y1 + exp(x0^2 + x1^2)"
    assert d.scalars() == [1.0, 1.0, 1.0, 1.0, 1.0]
    lin = BioData(os.path.join(INPUT, "linear.prm"))
    assert lin.has_solution and lin.constants["D_1"] == 4.0


def test_two_pressure_data():
    d = TwoPressureData(os.path.join(INPUT, "full_nonlinear.prm"))
    assert d.nonlinear and d.constants["A"] == 3.0
    assert d.scalars() == [1.0, 4.0, 2.0, 1.0]
    assert not TwoPressureData(os.path.join(INPUT, "linear_time.prm")).nonlinear


def test_malformed_lines(tmp_path):
    p = tmp_path / "c.prm"
    p.write_text("macro_rhs = 1\n")
    with pytest.raises(PrmError):
        EllipticData(str(p))
    ph = ParameterHandler()
    ph.declare_entry("a", "1")
    with pytest.raises(PrmError):
        ph.parse_input(str(tmp_path / "missing.prm"))


def test_every_committed_parameter_file_belongs_to_exactly_one_driver():
    """ParameterHandler aborts on undeclared keys (pde_data.cpp:8-23, :40-65, :110-126), so each .prm works with its own
    driver only: *map_test / elliptic_* -> solve_elliptic, linear / biocase* -> solve_biomath, *time* /
    full_nonlinear / ... -> solve_parabolic; parsed_mapping and ss_* belong to playground/transform.cpp."""
    import os

    from helpers import INPUT
    from dealiiscale_b200.prm import BioData, EllipticData, TwoPressureData

    owners = {}
    for f in sorted(os.listdir(INPUT)):
        if not f.endswith(".prm"):
            continue
        ok = []
        for cls in (EllipticData, BioData, TwoPressureData):
            try:
                cls(os.path.join(INPUT, f))
                ok.append(cls.__name__)
            except Exception:  # noqa: BLE001 (ParameterHandler errors of any kind)
                pass
        owners[f] = ok
    assert all(len(v) <= 1 for v in owners.values()), owners
    assert {f for f, v in owners.items() if not v} == {"parsed_mapping.prm", "ss_lin_equiv.prm", "ss_smt_equivalent.prm"}
    assert owners["map_test.prm"] == ["EllipticData"] and owners["biocase.prm"] == ["BioData"]
    assert owners["linear.prm"] == ["BioData"]  # BASELINE.json configs[0] pairs it with solve_elliptic: see SURVEY.md 8d
    assert owners["linear_time.prm"] == owners["full_nonlinear.prm"] == ["TwoPressureData"]
