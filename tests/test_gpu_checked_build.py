"""Memory-safety evidence for the resident CG kernels: the bounds-asserting build (csrc/checked.cuh,
dealiiscale_b200/_build/libmsgpu_checked.so).

compute-sanitizer is closed on the GPU pool, so the kernels that do their own index arithmetic over dynamic shared
memory, distributed shared memory and tensor memory - k_cg_stencil, k_cg_strip_tmem, k_cg_cluster - are compiled a
second time with every such access range-checked (a violation prints its source line and traps, which fails the launch
and this test).  Each case runs in a fresh process, once through the product library and once through the checked one,
on mesh sizes chosen to reach every tile shape / strip shape / cluster size the dispatcher can pick; the two runs must
also agree bit for bit (the checks change no arithmetic).  Reference call sites of the solves: src/bio-multiscale/
micro.cpp:309-314, src/elliptic-elliptic/micro.cpp:175-182."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import INPUT, ROOT

pytestmark = pytest.mark.gpu

RUNNER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from dealiiscale_b200 import capi
if sys.argv[2] == "checked":
    capi.LIB_PATH = os.path.join(sys.argv[1], "dealiiscale_b200", "_build", "libmsgpu_checked.so")
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import BioData, EllipticData
kind, prm, cells, n_sys, out = sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
rng = np.random.default_rng(11)
xy = rng.uniform(-0.9, 0.9, size=(n_sys, 2))
problem, cls = (capi.BIO, BioData) if kind == "bio" else (capi.ELLIPTIC, EllipticData)
ms = MicroSolver(problem, cls(os.path.join(sys.argv[1], "input", prm)), cells, mesh_kind=capi.MESH_SUBDIVIDED)
ms.max_it = 20000  # the elliptic driver's own limit (1000, micro.cpp:175) is for its coarser meshes
ms.set_grid_locations(xy)
ms.setup()
u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
w = np.cos(xy[:, 1]) * xy[:, 0] - 0.1
ms.set_macro_solution(u, w if kind == "bio" else None)
its = ms.run()
variant = ms.stats()["cg_variant"]
x = ms.solutions()
ms.zero_solutions()
its2 = ms.run()   # a second cold start: the class-table hint is known by now, only one CG kernel is launched
x2 = ms.solutions()
ms.close()
np.savez(out, x=x, its=its, x2=x2, its2=its2, variant=variant)
print("variant", variant)
"""

# kind, prm, cells, systems, env, CG variant the dispatcher must pick (kernels.h: 3 tensor-memory strips, 4 cluster,
# 5 class-compressed)
CASES = [
    ("bio", "biocase.prm", 64, 5, {}, 5),                                # 65 x 65: 3 x 6 tiles, shared row, dead column
    ("bio", "biocase.prm", 64, 3, {"MSGPU_STENCIL_VARIANT": "1"}, 5),    # the same with the generic masks
    ("bio", "biocase.prm", 64, 3, {"MSGPU_STENCIL_TMEM": "0"}, 5),       # vectors in registers
    ("bio", "linear.prm", 53, 4, {}, 5),                                 # 54 x 54 = 18 x 3 = 9 x 6
    ("elliptic", "map_test.prm", 64, 4, {}, 5),                          # Dirichlet block 63 x 63: 3 x 7 tiles
    ("elliptic", "lin_map_test.prm", 57, 4, {}, 5),                      # 56 x 56 block: dead tile column
    ("bio", "biocase.prm", 64, 3, {"MSGPU_CG_STENCIL": "0"}, 3),         # k_cg_strip_tmem<256, 17>
    ("elliptic", "elliptic_non_linear.prm", 64, 3, {}, 3),               # not class-uniform: general kernel behind the flag
    ("bio", "biocase.prm", 40, 6, {"MSGPU_CG_STENCIL": "0"}, 3),         # 41 x 41 = 1681 rows: <256, 9>, two CTAs per SM
    ("bio", "biocase.prm", 66, 3, {"MSGPU_CG_STENCIL": "0"}, 3),         # 67 x 67 = 4489 rows: <512, 9>
    ("bio", "biocase.prm", 96, 2, {"MSGPU_CG_STENCIL": "0"}, 4),         # 97 x 97 = 9409 rows: cluster of 3 CTAs
    ("elliptic", "elliptic_non_linear.prm", 96, 2, {}, 4),               # general matrices on a cluster of 3 CTAs
    ("elliptic", "elliptic_linear.prm", 128, 2, {}, 4),                  # 129 x 129 = 16641 rows: cluster of 4 CTAs
]


def _checked_lib():
    """The library normally travels with the tree (__graft_entry__.build()); nvcc is on the GPU boxes too."""
    lib = os.path.join(ROOT, "dealiiscale_b200", "_build", "libmsgpu_checked.so")
    if not os.path.exists(lib):
        sys.path.insert(0, ROOT)
        from dealiiscale_b200 import build as b
        b.build_checked()
    assert os.path.exists(lib)
    return lib


def _run(which, kind, prm, cells, n_sys, env, out):
    e = dict(os.environ)
    for k in ("MSGPU_CG_STENCIL", "MSGPU_STENCIL_TMEM", "MSGPU_STENCIL_VARIANT"):
        e.pop(k, None)
    e.update(env)
    res = subprocess.run([sys.executable, "-c", RUNNER, ROOT, which, kind, prm, str(cells), str(n_sys), out],
                         capture_output=True, text=True, env=e, timeout=600)
    assert res.returncode == 0, f"{which} library: rc {res.returncode}\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}"
    assert "checked build" not in res.stdout, res.stdout[-2000:]
    return np.load(out)


@pytest.mark.parametrize("kind,prm,cells,n_sys,env,variant", CASES)
def test_no_access_leaves_its_window(kind, prm, cells, n_sys, env, variant, tmp_path):
    lib = _checked_lib()
    chk = _run("checked", kind, prm, cells, n_sys, env, str(tmp_path / "checked.npz"))
    ref = _run("product", kind, prm, cells, n_sys, env, str(tmp_path / "product.npz"))
    assert int(ref["variant"]) & 0xff == variant or int(ref["variant"]) >> 8 == variant, int(ref["variant"])
    assert int(chk["variant"]) == int(ref["variant"])
    for key in ("its", "its2"):
        assert np.array_equal(chk[key], ref[key])
    for key in ("x", "x2"):
        assert np.array_equal(chk[key], ref[key])  # same arithmetic, same order: the checks are the only difference
    assert np.isfinite(ref["x"]).all() and ref["its"].min() > 0


def test_the_checks_fire(tmp_path):
    """The trap path itself: a checked build must fail loudly when a window is violated.  MSGPU_CHECKED_SELFTEST makes
    k_cg_stencil read one double in front of its shared-memory window."""
    lib = _checked_lib()
    e = dict(os.environ)
    e["MSGPU_CHECKED_SELFTEST"] = "1"
    res = subprocess.run([sys.executable, "-c", RUNNER, ROOT, "checked", "bio", "biocase.prm", "64", "2",
                          str(tmp_path / "x.npz")], capture_output=True, text=True, env=e, timeout=600)
    assert res.returncode != 0
    assert "[msgpu checked build]" in res.stdout + res.stderr
