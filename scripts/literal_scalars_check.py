"""How many CG iteration counts differ from the CPU oracle's, with the default scalar recurrence
(MSGPU_FAST_SCALARS=1: beta = |g'|^2 * (1 / |g|^2) with the reciprocal taken a step earlier, stop test on the squared
norm) and with the literal one of deal.II (-DMSGPU_FAST_SCALARS=0: res = sqrt(|g|^2), beta = res^2 / gh).

    python scripts/literal_scalars_check.py build     (CPU box: builds dealiiscale_b200/_build/libmsgpu_literal.so)
    python scripts/literal_scalars_check.py run        (GPU: prints one table per library)
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dealiiscale_b200 import build as b  # noqa: E402

LIB = os.path.join(b.BUILD, "libmsgpu_literal.so")


def build():
    objdir = os.path.join(b.BUILD, "obj")
    files, _ = b._sources()
    objs = []
    for src in files:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if os.path.basename(src) in ("kernels_cg.cu", "kernels_cg_stencil.cu"):
            obj = os.path.join(objdir, os.path.basename(src) + ".literal.o")
            subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-DMSGPU_FAST_SCALARS=0", "-c", src, "-o", obj], cwd=b.CSRC)
        objs.append(obj)
    subprocess.check_call([b._nvcc(), "-shared", "-o", LIB] + objs + ["-ldl"], cwd=b.CSRC)
    print("built", LIB)


def worker(lib):
    from dealiiscale_b200 import capi
    if lib:
        capi.LIB_PATH = lib
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData, EllipticData
    from helpers import INPUT, OracleBio, OracleElliptic

    rows = []
    bio = [("linear.prm", 4, 16), ("biocase.prm", 3, 8), ("biocase-a.prm", 2, 8), ("biocase-b.prm", 3, 8),
           ("biocase-c.prm", 3, 8), ("biocase-curved.prm", 3, 8), ("biocase-d.prm", 2, 16), ("biocase.prm", 4, 64),
           ("linear.prm", 3, 64), ("biocase-b.prm", 4, 47)]
    for prm, Mc, mc in bio:
        o = OracleBio(prm, Mc, mc, threads=8)
        xy = o.grid_locations()
        u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
        w = np.cos(xy[:, 1]) * xy[:, 0] - 0.1
        o.micro_run(u, w)
        ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, prm)), mc)
        ms.set_grid_locations(xy)
        ms.setup()
        ms.set_macro_solution(u, w)
        its = ms.run()
        d = np.abs(its - o.iters())
        big = o.iters() > o.n
        rows.append((f"bio {prm} {Mc} {mc}", int(o.N), int((d == 0).sum()), int(((d > 0) & ~big).sum()), int(d[~big].max(initial=0)),
                     int(big.sum()), int(ms.stats()["cg_variant"])))
        ms.close()
        o.close()
    for prm, Mr, mr in [("elliptic_linear.prm", 2, 4), ("elliptic_non_linear.prm", 2, 4), ("map_test.prm", 2, 5),
                        ("rot_map_test.prm", 1, 6), ("elliptic_non_linear.prm", 2, 5)]:
        o = OracleElliptic(prm, Mr, mr)
        xy = o.grid_locations()
        u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
        o.micro_run(u)
        ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, prm)), 2 ** mr)
        ms.set_grid_locations(xy)
        ms.setup()
        ms.set_macro_solution(u)
        its = ms.run()
        d = np.abs(its - o.iters())
        big = o.iters() > (2 ** mr - 1) ** 2
        rows.append((f"elliptic {prm} {Mr} {mr}", int(o.N), int((d == 0).sum()), int(((d > 0) & ~big).sum()),
                     int(d[~big].max(initial=0)), int(big.sum()), int(ms.stats()["cg_variant"])))
        ms.close()
        o.close()
    print(json.dumps(rows))


def run():
    for name, lib in (("default (MSGPU_FAST_SCALARS=1)", ""), ("literal (MSGPU_FAST_SCALARS=0)", LIB)):
        out = subprocess.run([sys.executable, __file__, "worker", lib], capture_output=True, text=True)
        if out.returncode != 0:
            print(name, "FAILED", out.stderr[-2000:])
            continue
        rows = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"--- {name}")
        print(f"{'case':44s} {'systems':>8s} {'equal':>6s} {'differ':>7s} {'max |d|':>8s} {'> n steps':>10s} {'kernel':>7s}")
        tot = [0, 0, 0, 0]
        for r in rows:
            print(f"{r[0]:44s} {r[1]:8d} {r[2]:6d} {r[3]:7d} {r[4]:8d} {r[5]:10d} {r[6]:7d}")
            tot = [tot[0] + r[1], tot[1] + r[2], tot[2] + r[3], max(tot[3], r[4])]
        print(f"{'total':44s} {tot[0]:8d} {tot[1]:6d} {tot[2]:7d} {tot[3]:8d}")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "worker":
        worker(sys.argv[2] if len(sys.argv) > 2 else "")
    else:
        run()
