"""Phase timings of the micro path at a given size (GPU box)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import EllipticData, BioData

prm = sys.argv[1] if len(sys.argv) > 1 else "elliptic_non_linear.prm"
macro_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 64
micro_cells = int(sys.argv[3]) if len(sys.argv) > 3 else 64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
max_it = int(sys.argv[5]) if len(sys.argv) > 5 else 0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "input", prm)
bio = prm.startswith("bio") or prm == "linear.prm"
data = BioData(path) if bio else EllipticData(path)
m = macro_cells + 1
g = np.linspace(-1, 1, m)
xy = np.array([(x, y) for y in g for x in g])
ms = MicroSolver(capi.BIO if bio else capi.ELLIPTIC, data, micro_cells, mesh_kind=capi.MESH_SUBDIVIDED)
if max_it:
    ms.max_it = max_it
ms.set_grid_locations(xy)
t0 = time.time(); ms.setup(); print("setup s %.3f" % (time.time() - t0))
u = np.sin(xy[:, 0]) * xy[:, 1]
ms.set_macro_solution(u, u if bio else None)
ms.enable_timing(True)
for r in range(reps):
    ms.zero_solutions()
    t0 = time.time()
    try:
        its = ms.run()
    except capi.MsgpuError as e:
        print("solve:", e); its = ms.last_iterations
    c = ms.coupling()
    dt = time.time() - t0
    st = ms.stats(reset=True)
    N, n = ms.num_grids, ms.n_dofs
    print(f"{prm} N={N} n={n} rep {r}: wall {dt*1e3:.2f} ms  DoF-solves/s {N*n/dt:.3e}  its mean {its.mean():.1f} max {its.max()}  "
          f"tables {st['ms_tables']:.2f} moments {st['ms_moments']:.2f} gather {st['ms_gather']:.2f} cg {st['ms_cg']:.2f} variant {st['cg_variant']}")
