"""Probe: k_cg_cluster on systems that are NOT class-uniform (elliptic_non_linear.prm) at 96 / 128 cells, against the
streaming fallback (MSGPU_CG_CLUSTER=0).  Usage (GPU): python scripts/cluster_nonuniform_probe.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[1])
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import EllipticData
cells = int(sys.argv[2])
rng = np.random.default_rng(11)
xy = rng.uniform(-0.9, 0.9, size=(2, 2))
ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(sys.argv[1], "input", "elliptic_non_linear.prm")), cells, mesh_kind=capi.MESH_SUBDIVIDED)
ms.max_it = int(sys.argv[3])
ms.set_grid_locations(xy); ms.setup()
ms.set_macro_solution(np.sin(xy[:, 0]) * xy[:, 1] + 0.3)
try:
    its = ms.run()
    print("cells", cells, "variant", ms.stats()["cg_variant"], "its", its, "x norm", float(np.abs(ms.solutions()).max()))
except Exception as e:
    print("cells", cells, "variant", ms.stats()["cg_variant"], "FAILED", e)
"""
for cells in (96, 128):
    for env in ({}, {"MSGPU_CG_CLUSTER": "0"}):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", CODE, ROOT, str(cells), "100000"], env=e, capture_output=True, text=True)
        print(env, r.stdout.strip()[-400:], r.stderr.strip()[-300:])
