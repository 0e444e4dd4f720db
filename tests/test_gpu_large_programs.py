"""Expressions with many live temporaries: the interpreter kernels keep one VM register file per thread in
dynamic shared memory and must opt in above the 48 KB default; a program that does not fit an SM is refused with a clear
message instead of a failed launch.  The reference's muparser path has no such limit
(src/tools/multiscale_function_parser.cpp:338-436)."""
import os

import numpy as np
import pytest

from helpers import INPUT

pytestmark = pytest.mark.gpu


def _many_temporaries(n, var="y0"):
    """Every t_i is used twice, so all n stay live between the sum and the product."""
    terms = [f"sin({i + 1}*{var}+y1)" for i in range(n)]
    return "(" + "+".join(terms) + ")+1e-30*(" + "*".join(terms) + ")"


@pytest.mark.parametrize("jit", ["1", "0"])
def test_sixty_live_temporaries_assemble_and_solve(jit, monkeypatch):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    monkeypatch.setenv("MSGPU_JIT", jit)
    data = EllipticData(os.path.join(INPUT, "elliptic_non_linear.prm"))
    rhs = _many_temporaries(60)
    xy = np.array([[0.1, 0.2], [-0.3, 0.4]])
    ms = MicroSolver(capi.ELLIPTIC, data, 8)
    ms.set_function(capi.FN_RHS, rhs, data.constants)
    ms.set_function(capi.FN_BC, _many_temporaries(55, "y1"), data.constants)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_macro_solution(np.zeros(2))
    ms.run()
    b = ms.righthandsides()
    assert np.isfinite(b).all() and np.abs(b).max() > 0
    # the same load from a mathematically identical expression that needs few temporaries at a time
    few = "+".join(f"sin({i + 1}*y0+y1)" for i in range(60))
    ref = MicroSolver(capi.ELLIPTIC, data, 8)
    ref.set_function(capi.FN_RHS, few, data.constants)
    ref.set_function(capi.FN_BC, "+".join(f"sin({i + 1}*y1+y1)" for i in range(55)), data.constants)
    ref.set_grid_locations(xy)
    ref.setup()
    ref.set_macro_solution(np.zeros(2))
    ref.run()
    assert np.abs(b - ref.righthandsides()).max() <= 1e-12 * np.abs(b).max()
    assert np.abs(ms.solutions() - ref.solutions()).max() <= 1e-10 * np.abs(ref.solutions()).max()
    ms.close()
    ref.close()


def test_a_program_that_does_not_fit_the_device_vm_is_refused_with_a_message(monkeypatch):
    """Hundreds of live temporaries: refused when the program is built (VM code length) or when it is launched (shared
    memory of an SM) - never a failed launch with 'invalid argument'."""
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    monkeypatch.setenv("MSGPU_JIT", "0")
    data = EllipticData(os.path.join(INPUT, "elliptic_non_linear.prm"))
    ms = MicroSolver(capi.ELLIPTIC, data, 8)
    # 18 x 18 = 324 shared point-level subexpressions from 18 distinct constants (the VM holds few constants)
    terms = [f"sin({i + 1}*y0+{j + 1}*y1)" for i in range(18) for j in range(18)]
    ms.set_function(capi.FN_BC, "(" + "+".join(terms) + ")+1e-30*(" + "*".join(terms) + ")", data.constants)
    ms.set_grid_locations(np.array([[0.1, 0.2]]))
    with pytest.raises(capi.MsgpuError) as err:
        ms.setup()
    assert err.value.code in (capi.ERR_UNSUPPORTED, capi.ERR_PARSE)
    assert "live temporaries" in str(err.value) or "too long for the device VM" in str(err.value)
    ms.close()
