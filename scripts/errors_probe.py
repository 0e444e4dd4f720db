"""Time of the per-grid error integrals (msgpu_errors) at the bench size."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import BioData
prm = sys.argv[1] if len(sys.argv) > 1 else "linear.prm"
macro = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = np.linspace(-1, 1, macro + 1)
xy = np.array([(a, b) for b in g for a in g])
ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", prm)), 64)
ms.set_grid_locations(xy); ms.setup()
ms.set_macro_solution(xy[:, 0], xy[:, 1]); ms.run()
for rep in range(3):
    ms.synchronize(); t0 = time.perf_counter()
    l2, h1 = ms.grid_errors()
    ms.synchronize(); print("grid_errors %.2f ms  (jit kernels %d)" % ((time.perf_counter() - t0) * 1e3, ms.stats()["jit_kernels"]), flush=True)
