"""A/B of the class-compressed CG kernel (kernels_cg_stencil.cu) against the general kernels on a B200:
same assembled systems, MSGPU_CG_STENCIL=0 vs default; prints the kernel variant, the largest relative
difference of the solutions, the iteration counts and the CG times.

    python scripts/stencil_check.py [quick]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dealiiscale_b200 import capi  # noqa: E402
from dealiiscale_b200.micro import MicroSolver  # noqa: E402
from dealiiscale_b200.prm import BioData, EllipticData, TwoPressureData  # noqa: E402

INPUT = os.path.join(ROOT, "input")


def lattice(n_side):
    g = np.linspace(-1.0, 1.0, n_side)
    X, Y = np.meshgrid(g, g, indexing="ij")
    return np.stack([X.ravel(), Y.ravel()], axis=1)


def run_case(problem, data, cells, n_side, label, steps=1):
    xy = lattice(n_side)
    u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
    w = np.cos(xy[:, 1]) * xy[:, 0] - 0.1
    out = {}
    for mode in ("0", "1"):
        os.environ["MSGPU_CG_STENCIL"] = mode
        ms = MicroSolver(problem, data, cells)
        ms.set_grid_locations(xy)
        ms.setup()
        ms.enable_timing(True)
        if problem == capi.PARABOLIC:
            ms.set_macro_solution(u)
            its = None
            ms.stats(reset=True)
            t0 = time.perf_counter()
            for i in range(steps):
                its = ms.iterate(0.01, 0.01 * (i + 1))
            wall = time.perf_counter() - t0
        else:
            ms.set_macro_solution(u, w if problem == capi.BIO else None)
            ms.run()  # warm-up (module load, hint discovery)
            ms.zero_solutions()
            ms.stats(reset=True)
            t0 = time.perf_counter()
            for i in range(steps):
                if problem == capi.BIO:
                    ms.zero_solutions()
                its = ms.run()
            wall = time.perf_counter() - t0
        st = ms.stats()
        out[mode] = (ms.solutions(), its.copy(), st["cg_variant"], st["ms_cg"] / steps, wall / steps)
        ms.close()
    x0, i0, v0, t0_, _ = out["0"]
    x1, i1, v1, t1_, _ = out["1"]
    num = np.linalg.norm(x0 - x1, axis=1)
    den = np.maximum(np.linalg.norm(x0, axis=1), 1e-300)
    print(f"{label:46s} N={len(xy):5d} n={x0.shape[1]:5d}  variant {v0}->{v1}  max rel diff {np.max(num / den):.2e}  "
          f"its mean {i0.mean():.1f}->{i1.mean():.1f} max|d| {np.abs(i0 - i1).max()}  cg ms {t0_:.3f}->{t1_:.3f}",
          flush=True)
    return np.max(num / den), np.abs(i0 - i1).max()


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    bad = 0
    cases = [
        (capi.BIO, BioData(os.path.join(INPUT, "biocase.prm")), 64, 9, "bio biocase 64 cells"),
        (capi.BIO, BioData(os.path.join(INPUT, "biocase-curved.prm")), 64, 5, "bio biocase-curved 64 cells"),
        (capi.BIO, BioData(os.path.join(INPUT, "linear.prm")), 32, 5, "bio linear 32 cells"),
        (capi.BIO, BioData(os.path.join(INPUT, "biocase-c.prm")), 40, 5, "bio biocase-c 40 cells"),
        (capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "map_test.prm")), 64, 5, "elliptic map_test 64 cells"),
        (capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "rot_map_test.prm")), 32, 5, "elliptic rot_map_test 32 cells"),
        (capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "elliptic_non_linear.prm")), 32, 5,
         "elliptic non_linear 32 cells (not uniform)"),
        (capi.PARABOLIC, TwoPressureData(os.path.join(INPUT, "linear_time.prm")), 64, 9, "parabolic linear_time 64"),
        (capi.PARABOLIC, TwoPressureData(os.path.join(INPUT, "linear_time.prm")), 45, 5, "parabolic linear_time 45"),
        (capi.PARABOLIC, TwoPressureData(os.path.join(INPUT, "full_nonlinear.prm")), 32, 5, "parabolic full_nonlinear 32"),
    ]
    for prob, data, cells, side, label in cases:
        d, di = run_case(prob, data, cells, side, label, steps=2 if prob == capi.PARABOLIC else 1)
        if not (d < 1e-9):
            bad += 1
            print("   ^^^ MISMATCH")
    if not quick:
        run_case(capi.BIO, BioData(os.path.join(INPUT, "biocase.prm")), 64, 65, "bio biocase 64 cells, 4225 systems", steps=3)
        run_case(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "map_test.prm")), 64, 65,
                 "elliptic map_test 64 cells, 4225 systems", steps=3)
        run_case(capi.PARABOLIC, TwoPressureData(os.path.join(INPUT, "linear_time.prm")), 64, 65,
                 "parabolic linear_time 64, 4225 systems", steps=3)
    print("stencil_check:", "FAILED" if bad else "ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
