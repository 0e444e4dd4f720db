"""One workload for profiling the class-compressed CG kernel: bio biocase.prm, 64 x 64-cell micro systems.

    python scripts/stencil_profile.py <n_systems> [prm] [cells]
prints ms per CG launch (CUDA events inside libmsgpu) for the default launch and for one CTA per SM.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dealiiscale_b200 import capi  # noqa: E402
from dealiiscale_b200.micro import MicroSolver  # noqa: E402
from dealiiscale_b200.prm import BioData  # noqa: E402


def main():
    n_sys = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    prm = sys.argv[2] if len(sys.argv) > 2 else "biocase.prm"
    cells = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    rng = np.random.default_rng(7)
    xy = rng.uniform(-1.0, 1.0, size=(n_sys, 2))
    u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
    w = np.cos(xy[:, 1]) * xy[:, 0] - 0.1
    ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", prm)), cells)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.enable_timing(True)
    ms.set_macro_solution(u, w)
    ms.run()
    for label, env in (("default", None), ("one CTA per SM", "1")):
        if env is None:
            os.environ.pop("MSGPU_STENCIL_CTAS", None)
        else:
            os.environ["MSGPU_STENCIL_CTAS"] = env
        ms.reload_switches()
        ms.stats(reset=True)
        reps = 3
        for _ in range(reps):
            ms.zero_solutions()
            its = ms.run()
        st = ms.stats()
        print(f"{label:16s} N={n_sys} variant {st['cg_variant']} its mean {its.mean():.1f}  cg {st['ms_cg'] / reps:.3f} ms"
              f"  ({st['ms_cg'] / reps * 1e3 / max(1, its.sum()) * 148:.3f} us per step and SM)", flush=True)
    os.environ.pop("MSGPU_STENCIL_CTAS", None)
    ms.close()


if __name__ == "__main__":
    main()
