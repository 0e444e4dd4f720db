"""A/B of code-generation variants of k_cg_stencil on the bench workload (4225 bio systems of 64 x 64 cells):
MSGPU_STENCIL_VARIANT = 0 (default) | 1 (masks on every row / column that could be shared or dead) | 2 (x updated
between the two reductions of a step instead of behind the second one).

    python scripts/stencil_variants.py [variants ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealiiscale_b200 import capi  # noqa: E402
from dealiiscale_b200.micro import MicroSolver  # noqa: E402
from dealiiscale_b200.prm import BioData  # noqa: E402

variants = sys.argv[1:] or ["0", "1", "2", "3"]
g = np.linspace(-1.0, 1.0, 65)
X, Y = np.meshgrid(g, g, indexing="ij")
xy = np.stack([X.ravel(), Y.ravel()], axis=1)
u, w = xy[:, 1] * np.sin(xy[:, 0]), xy[:, 0] * np.cos(xy[:, 1])
ref = None
for v in variants:
    os.environ["MSGPU_STENCIL_VARIANT"] = v
    ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", "biocase.prm")), 64)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.enable_timing(True)
    ms.set_macro_solution(u, w)
    for _ in range(2):
        ms.zero_solutions()
        ms.run()
    ms.stats(reset=True)
    reps = 5
    for _ in range(reps):
        ms.zero_solutions()
        its = ms.run()
    st = ms.stats()
    x = ms.solutions(0, 64)
    if ref is None:
        ref = x
    print(f"variant {v}: cg {st['ms_cg'] / reps:.3f} ms  its {its.mean():.2f}  kernel {st['cg_variant']}  "
          f"max diff to first variant {np.abs(x - ref).max():.2e}", flush=True)
    ms.close()
