"""`.prm` parameter files with deal.II ``ParameterHandler`` semantics, and the three data classes
of the reference that turn them into named expressions.

Mirrors ``src/tools/pde_data.cpp`` of the reference:

* ``MultiscaleData``  (:7-36)    -> :class:`EllipticData`   (``solve_elliptic``)
* ``BioData``         (:38-108)  -> :class:`BioData`        (``solve_biomath``)
* ``TwoPressureData`` (:110-159) -> :class:`TwoPressureData` (``solve_parabolic``)

Same behaviour as the reference: every key is declared with a default first, the file overrides,
an undeclared key aborts the parse (so each file only works with its own driver), ``#`` starts a
comment, physical constants are declared as numbers and injected into every expression as
constants.  Only parsing and bookkeeping happen here; expressions are evaluated on the device.
"""
from __future__ import annotations

from collections import OrderedDict


class PrmError(ValueError):
    """Raised where the reference's ParameterHandler / muParser would throw."""


class ParameterHandler:
    def __init__(self):
        self._values: "OrderedDict[str, str]" = OrderedDict()

    def declare_entry(self, key: str, default: str) -> None:
        self._values[key] = default

    def parse_input(self, path: str) -> None:
        try:
            with open(path, "r") as fh:
                lines = fh.readlines()
        except OSError as exc:
            raise PrmError(f"cannot open parameter file {path}: {exc}") from exc
        for lineno, raw in enumerate(lines, 1):
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if not (line.startswith("set ") or line.startswith("set\t")):
                raise PrmError(f"{path}:{lineno}: expected 'set key = value'")
            if "=" not in line:
                raise PrmError(f"{path}:{lineno}: missing '='")
            key, value = line[3:].split("=", 1)
            key, value = key.strip(), value.strip()
            if key not in self._values:
                raise PrmError(f"{path}:{lineno}: no entry with name <{key}> was declared")
            self._values[key] = value

    def get(self, key: str) -> str:
        return self._values[key]

    def get_double(self, key: str) -> float:
        return float(self._values[key])

    def get_bool(self, key: str) -> bool:
        return self._values[key] in ("true", "yes", "on")


_MAP_DEFAULT = "(3+x0+x1)*(1.6*y0 - 1.6*y1);(3+x0+x1)*(2.4*y0 + 2.4*y1);"
_JAC_DEFAULT = "1.6*(3+x0+x1);-1.6*(3+x0+x1);2.4*(3+x0+x1);2.4*(3+x0+x1);"


class EllipticData:
    """``MultiscaleData<2>`` (pde_data.cpp:7-36).  No constants."""

    def __init__(self, param_file: str):
        p = self.params = ParameterHandler()
        p.declare_entry("macro_geometry", "[-1,1]x[-1,1]")
        p.declare_entry("macro_rhs", " x0^2*sin(x0*x1) + x1^2*sin(x0*x1) + 2*cos(x0 + x1)")
        p.declare_entry("macro_solution", "sin(x0*x1) + cos(x0 + x1)")
        p.declare_entry("macro_bc", "sin(x0*x1) + cos(x0 + x1)")
        p.declare_entry("micro_geometry", "[-1,1]x[-1,1]")
        p.declare_entry("micro_rhs", "-sin(x0*x1) - cos(x0 + x1)")
        p.declare_entry("micro_solution", "y0*y1 + exp(x0^2 + x1^2)")
        p.declare_entry("micro_bc", "y0*y1 + exp(x0^2 + x1^2)")
        p.declare_entry("mapping", _MAP_DEFAULT)
        p.declare_entry("jac_mapping", _JAC_DEFAULT)
        p.parse_input(param_file)
        self.constants: dict = {}
        self.macro = {k: p.get("macro_" + k) for k in ("rhs", "bc", "solution")}
        self.micro = {
            "mapping": p.get("mapping"),
            "jac_mapping": p.get("jac_mapping"),
            "rhs": p.get("micro_rhs"),
            "bc": p.get("micro_bc"),
            "solution": p.get("micro_solution"),
        }


class BioData:
    """``BioData<2>`` (pde_data.cpp:38-108)."""

    CONSTANT_DEFAULTS = OrderedDict(
        [("D_1", 2.0), ("D_2", 2.0), ("kappa_1", 1.0), ("kappa_2", 1.0), ("kappa_3", 1.0), ("kappa_4", 1.0)]
    )

    def __init__(self, param_file: str):
        p = self.params = ParameterHandler()
        macro_default = "sin(x0*x1) + cos(x0 + x1)"
        macro_rhs_default = " x0^2*sin(x0*x1) + x1^2*sin(x0*x1) + 2*cos(x0 + x1)"
        micro_default = "y0*y1 + exp(x0^2 + x1^2)"
        p.declare_entry("macro_geometry", "[-1,1]x[-1,1]")
        p.declare_entry("has_solution", "true")
        p.declare_entry("solution_u", macro_default)
        p.declare_entry("bulk_rhs_u", macro_rhs_default)
        p.declare_entry("bc_u_1", macro_default)
        p.declare_entry("bc_u_2", macro_default)
        p.declare_entry("inflow_measure", "x0 + x1")
        p.declare_entry("outflow_measure", "x0 + x1")
        p.declare_entry("solution_w", macro_default)
        p.declare_entry("bulk_rhs_w", macro_rhs_default)
        p.declare_entry("bc_w_1", macro_default)
        p.declare_entry("bc_w_2", macro_default)
        p.declare_entry("micro_geometry", "[-1,1]x[-1,1]")
        p.declare_entry("solution_v", "-D_1 * sin(x0*x1) - cos(x0 + x1)")
        p.declare_entry("bulk_rhs_v", micro_default)
        for i in (1, 2, 3, 4):
            p.declare_entry(f"bc_v_{i}", micro_default)
        p.declare_entry("mapping", _MAP_DEFAULT)
        p.declare_entry("jac_mapping", _JAC_DEFAULT)
        p.declare_entry("num_threads", "-1")
        for name, val in self.CONSTANT_DEFAULTS.items():
            p.declare_entry(name, f"{val:f}")  # std::to_string(double)
        p.parse_input(param_file)
        self.constants = {name: p.get_double(name) for name in self.CONSTANT_DEFAULTS}
        self.has_solution = p.get_bool("has_solution")
        self.macro = {
            k: p.get(k)
            for k in (
                "solution_u", "bulk_rhs_u", "bc_u_1", "bc_u_2", "inflow_measure", "outflow_measure",
                "solution_w", "bulk_rhs_w", "bc_w_1", "bc_w_2",
            )
        }
        self.micro = {
            "mapping": p.get("mapping"),
            "jac_mapping": p.get("jac_mapping"),
            "rhs": p.get("bulk_rhs_v"),
            "solution": p.get("solution_v"),
            "bc_v_1": p.get("bc_v_1"),
            "bc_v_2": p.get("bc_v_2"),
            "bc_v_3": p.get("bc_v_3"),
            "bc_v_4": p.get("bc_v_4"),
        }

    def scalars(self):
        """kappa_1..4, D_2 in the order msgpu_set_scalars expects."""
        c = self.constants
        return [c["kappa_1"], c["kappa_2"], c["kappa_3"], c["kappa_4"], c["D_2"]]


class TwoPressureData:
    """``TwoPressureData<2>`` (pde_data.cpp:110-171); every function but ``init_rho`` sees ``t``."""

    CONSTANT_DEFAULTS = OrderedDict([("A", 0.8), ("D", 1.0), ("theta", 0.5), ("kappa", 1.0), ("p_F", 4.0), ("R", 2.0)])

    def __init__(self, param_file: str):
        p = self.params = ParameterHandler()
        p.declare_entry("macro_geometry", "[-1,1]x[-1,1]")
        p.declare_entry(
            "macro_rhs",
            "-A*(2*(2*x0^2 + 1)*exp(t^2 + x0^2 + x1^2) + 2*(2*x1^2 + 1)*exp(t^2 + x0^2 + x1^2)) - 12*theta*exp(t^2 + x0^2 + x1^2)",
        )
        p.declare_entry("macro_solution", "exp(t^2 + x0^2 + x1^2)")
        p.declare_entry("macro_bc", "exp(t^2 + x0^2 + x1^2)")
        p.declare_entry("micro_geometry", "[-1,1]x[-1,1]")
        p.declare_entry("micro_rhs", "-D*(2*(sin(y0)^2 - cos(y0)^2) + 2*(-sin(y1)^2 + cos(y1)^2))")
        p.declare_entry("micro_solution", "sin(y1)^2 + cos(y0)^2 + 2")
        p.declare_entry("micro_bc_neumann", "-2*D*y0*sin(y0)*cos(y0)")
        p.declare_entry(
            "micro_bc_robin",
            "2*D*y1*sin(y1)*cos(y1) - kappa*(-R*(sin(y1)^2 + cos(y0)^2 + 2) + p_F + exp(t^2 + x0^2 + x1^2))",
        )
        p.declare_entry("init_rho", "sin(y1)^2 + cos(y0)^2 + 2")
        p.declare_entry("nonlinear", "false")
        for name, val in self.CONSTANT_DEFAULTS.items():
            p.declare_entry(name, f"{val:f}")
        p.parse_input(param_file)
        self.constants = {name: p.get_double(name) for name in self.CONSTANT_DEFAULTS}
        self.nonlinear = p.get_bool("nonlinear")
        self.macro = {k: p.get("macro_" + k) for k in ("rhs", "bc", "solution")}
        self.micro = {
            "rhs": p.get("micro_rhs"),
            "solution": p.get("micro_solution"),
            "bc_neumann": p.get("micro_bc_neumann"),
            "bc_robin": p.get("micro_bc_robin"),
            "init_rho": p.get("init_rho"),
        }

    def scalars(self):
        """kappa, p_F, R, D in the order msgpu_set_scalars expects."""
        c = self.constants
        return [c["kappa"], c["p_F"], c["R"], c["D"]]
