"""Cycle stamps of the class-compressed CG kernel (debug build, -DMSGPU_STENCIL_STAMPS): mean clock64 cycles per CG
step and phase of CTA 0, read back through the residual array.  Builds its own copy of the library next to the
product's (dealiiscale_b200/_build/libmsgpu_stamps.so) so that the product is never a debug build.

    python scripts/stencil_stamps.py build          (on the CPU box)
    python scripts/stencil_stamps.py run [n_sys]    (on the GPU)
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealiiscale_b200 import build as b  # noqa: E402

LIB = os.path.join(b.BUILD, "libmsgpu_stamps.so")
PHASES = ["closing barrier", "SpMV + part of d.h", "first reduction", "alpha", "x, g updates + part of g.g",
          "second reduction", "beta, direction, publish"]


def build():
    objdir = os.path.join(b.BUILD, "obj")
    files, _ = b._sources()
    objs = []
    for src in files:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if src.endswith("kernels_cg_stencil.cu"):
            obj = os.path.join(objdir, "kernels_cg_stencil.stamps.o")
            subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-DMSGPU_STENCIL_STAMPS", "-c", src, "-o", obj], cwd=b.CSRC)
        objs.append(obj)
    subprocess.check_call([b._nvcc(), "-shared", "-o", LIB] + objs + ["-ldl"], cwd=b.CSRC)
    print("built", LIB)


def run(n_sys):
    from dealiiscale_b200 import capi
    capi.LIB_PATH = LIB
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    rng = np.random.default_rng(7)
    xy = rng.uniform(-1.0, 1.0, size=(n_sys, 2))
    ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", "biocase.prm")), 64)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.enable_timing(True)
    ms.set_macro_solution(np.sin(xy[:, 0]) * xy[:, 1] + 0.3, np.cos(xy[:, 1]) * xy[:, 0] - 0.1)
    ms.run()
    ms.zero_solutions()
    ms.stats(reset=True)
    ms.assemble_system()
    its = ms.solve(want_residuals=True)
    st = ms.stats()
    cyc = ms.last_residuals[:7]
    print(f"N={n_sys} variant {st['cg_variant']} its {its.mean():.1f} cg {st['ms_cg']:.3f} ms")
    for name, c in zip(PHASES, cyc):
        print(f"  {name:32s} {c:8.1f} cycles")
    print(f"  {'sum':32s} {cyc.sum():8.1f} cycles per CG step")
    ms.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 296)
