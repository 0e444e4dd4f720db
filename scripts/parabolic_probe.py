"""Phase timings of RhoSolver::iterate at the finest level of solve_parabolic (h_inv = 64: 4225 grids x 4225 DoFs)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import TwoPressureData

prm = sys.argv[1] if len(sys.argv) > 1 else "linear_time.prm"
h_inv = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = np.linspace(-1, 1, h_inv + 1)
xy = np.array([(a, b) for b in g for a in g])
ms = MicroSolver(capi.PARABOLIC, TwoPressureData(os.path.join(ROOT, "input", prm)), h_inv)
ms.set_grid_locations(xy)
t0 = time.time(); ms.setup(); print(f"setup {time.time() - t0:.2f} s")
pi = np.cos(xy[:, 0]) * xy[:, 1] + 2
ms.set_macro_solution(pi)
dt = 1.0 / 256
for w in range(3):
    ms.iterate(dt, dt * (w + 1))
ms.enable_timing(True); ms.stats(reset=True)
t0 = time.time()
K = 10
for s in range(K):
    its = ms.iterate(dt, dt * (s + 4))
    ms.coupling()
wall = (time.time() - t0) / K
st = ms.stats(reset=True)
print(f"{prm} h_inv {h_inv}: wall {wall * 1e3:.2f} ms per step; CG steps mean {its.mean():.1f} max {its.max()}; variant {st['cg_variant']}")
print({k: round(st[k] / K, 3) for k in ("ms_tables", "ms_moments", "ms_gather", "ms_cg", "ms_coupling")}, "launches/step", st["launches"] / K)
