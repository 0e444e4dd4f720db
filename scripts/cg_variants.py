"""A/B of the CG kernel variants at the bench size: bitwise comparison of the solutions + timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import BioData

prm = sys.argv[1] if len(sys.argv) > 1 else "biocase-curved.prm"
macro = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g = np.linspace(-1, 1, macro + 1)
xy = np.array([(a, b) for b in g for a in g])
u, w = xy[:, 1] * np.sin(xy[:, 0]), xy[:, 0] * np.cos(xy[:, 1])
ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", prm)), 64)
ms.set_grid_locations(xy); ms.setup(); ms.set_macro_solution(u, w); ms.assemble_system()
results = {}
for name, env in (("strip", {"MSGPU_CG_TMEM": "0"}), ("tmem", {}), ("tmem512", {"MSGPU_CG_TMEM": "5"}), ("strided", {"MSGPU_CG_STRIDED": "1"})):
    for k in ("MSGPU_CG_TMEM", "MSGPU_CG_STRIDED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ms.enable_timing(True)
    for rep in range(3):
        ms.zero_solutions(); ms.stats(reset=True)
        its = ms.solve()
        st = ms.stats(reset=True)
    results[name] = (ms.solutions(), its.copy(), st["ms_cg"], st["cg_variant"])
    print(f"{name}: variant {st['cg_variant']} cg {st['ms_cg']:.3f} ms  its mean {its.mean():.1f}", flush=True)
ref = results["strip"]
for name in ("tmem", "tmem512", "strided"):
    x, its, _, _ = results[name]
    print(name, "bitwise equal to strip:", np.array_equal(x, ref[0]), "max rel diff %.2e" % (np.abs(x - ref[0]).max() / np.abs(ref[0]).max()), "its equal:", np.array_equal(its, ref[1]))
