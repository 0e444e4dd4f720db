"""Cluster CG (one system across 2-8 SMs) against the streaming fallback kernel, sizes m > 67."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import BioData

prm = sys.argv[1] if len(sys.argv) > 1 else "biocase.prm"
macro = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for cells in [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "68,80,96,128").split(",")]:
    g = np.linspace(-1, 1, macro + 1)
    xy = np.array([(a, b) for b in g for a in g])
    u, w = xy[:, 1] * np.sin(xy[:, 0]), xy[:, 0] * np.cos(xy[:, 1])
    ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", prm)), cells)
    ms.set_grid_locations(xy); ms.setup(); ms.set_macro_solution(u, w); ms.assemble_system()
    out = {}
    for name, env in (("cluster", {}), ("global", {"MSGPU_CG_VARIANT": "global"})):
        os.environ.pop("MSGPU_CG_VARIANT", None)
        os.environ.update(env)
        ms.enable_timing(True)
        for rep in range(2):
            ms.zero_solutions(); ms.stats(reset=True)
            its = ms.solve()
            st = ms.stats(reset=True)
        out[name] = (ms.solutions(), its.copy())
        print(f"cells {cells} n {ms.n_dofs} N {len(xy)} {name}: variant {st['cg_variant']} cg {st['ms_cg']:.3f} ms its mean {its.mean():.1f}", flush=True)
    os.environ.pop("MSGPU_CG_VARIANT", None)
    d = np.abs(out["cluster"][0] - out["global"][0]).max() / np.abs(out["global"][0]).max()
    print(f"   max rel diff cluster vs global {d:.2e}  its diff max {np.abs(out['cluster'][1] - out['global'][1]).max()}", flush=True)
    ms.close()
