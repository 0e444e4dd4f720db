"""Profiling workload with ONE launch configuration (see stencil_profile.py): n_systems bio systems of 64 x 64 cells,
CTAs per SM taken from MSGPU_STENCIL_CTAS in the environment."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dealiiscale_b200 import capi  # noqa: E402
from dealiiscale_b200.micro import MicroSolver  # noqa: E402
from dealiiscale_b200.prm import BioData  # noqa: E402

n_sys = int(sys.argv[1]) if len(sys.argv) > 1 else 592
rng = np.random.default_rng(7)
xy = rng.uniform(-1.0, 1.0, size=(n_sys, 2))
ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", "biocase.prm")), 64)
ms.set_grid_locations(xy)
ms.setup()
ms.enable_timing(True)
ms.set_macro_solution(np.sin(xy[:, 0]) * xy[:, 1] + 0.3, np.cos(xy[:, 1]) * xy[:, 0] - 0.1)
ms.run()
ms.stats(reset=True)
for _ in range(2):
    ms.zero_solutions()
    its = ms.run()
st = ms.stats()
print(f"N={n_sys} CTAs/SM={os.environ.get('MSGPU_STENCIL_CTAS', 'max')} variant {st['cg_variant']} its {its.mean():.1f} cg {st['ms_cg'] / 2:.3f} ms")
ms.close()
