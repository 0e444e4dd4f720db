"""Small end-to-end case of all three micro problems for compute-sanitizer (GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import BioData, EllipticData, TwoPressureData

inp = os.path.join(ROOT, "input")
g = np.linspace(-1, 1, 3)
xy = np.array([(a, b) for b in g for a in g])
for cells in (8, 16):
    ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(inp, "elliptic_non_linear.prm")), cells)
    ms.set_grid_locations(xy); ms.setup(); ms.set_macro_solution(np.sin(xy[:, 0])); ms.run(); ms.coupling(); ms.grid_errors(); ms.close()
    mb = MicroSolver(capi.BIO, BioData(os.path.join(inp, "linear.prm")), cells + 1)
    mb.set_grid_locations(xy); mb.setup(); mb.set_macro_solution(xy[:, 0], xy[:, 1]); mb.run(); mb.run(); mb.coupling(); mb.grid_errors(); mb.grid_residuals(); mb.close()
    mp_ = MicroSolver(capi.PARABOLIC, TwoPressureData(os.path.join(inp, "full_nonlinear.prm")), cells + 3)
    mp_.set_grid_locations(xy); mp_.setup(); mp_.set_macro_solution(1 + xy[:, 0]); mp_.iterate(0.25, 0.25); mp_.iterate(0.25, 0.5); mp_.coupling(); mp_.grid_errors(); mp_.close()
os.environ["MSGPU_CG_VARIANT"] = "global"
ms = MicroSolver(capi.BIO, BioData(os.path.join(inp, "biocase-b.prm")), 12)
ms.set_grid_locations(xy); ms.setup(); ms.set_macro_solution(xy[:, 0], xy[:, 1]); ms.run(); ms.close()
print("sanitize case done")
