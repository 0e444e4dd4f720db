"""Where two CG runs that both stop at |g| <= 1e-12 may differ by more than 1e-10.

On the kappa_2 = kappa_4 = 0.01 set (biocase-c) and the curved-map set the micro matrices are so ill-conditioned that
the stopping rule of the reference (absolute recursive residual 1e-12, src/bio-multiscale/micro.cpp:309-314) leaves an
error of up to cond(A) * 1e-12 / |A| in the iterate; WHICH vector inside that ball a run ends on is decided by the
rounding of its dot products.  The GPU and the CPU oracle sum in different orders, so they may differ by twice that
radius - and by no more.  This test measures the radius instead of asserting it: both iterates are compared with a
DIRECT solve of the oracle's system, and the GPU's distance to it must not exceed twice the oracle's own (plus the
rounding of the direct solve).  The numbers are printed (pytest -s) and kept in the assertion messages."""
import os

import numpy as np
import pytest

from helpers import INPUT, OracleBio, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prm,macro_cells,micro_cells", [("biocase-c.prm", 3, 8), ("biocase-curved.prm", 3, 8),
                                                         ("biocase-c.prm", 2, 16), ("biocase-curved.prm", 2, 16)])
def test_gpu_iterate_is_as_close_to_the_direct_solution_as_the_oracle_iterate(prm, macro_cells, micro_cells):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    o = OracleBio(prm, macro_cells, micro_cells)
    xy = o.grid_locations()
    u = np.sin(xy[:, 0]) * xy[:, 1] + 0.3
    w = np.cos(xy[:, 1]) * xy[:, 0] - 0.1
    o.micro_run(u, w)
    ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, prm)), micro_cells)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_macro_solution(u, w)
    ms.assemble_and_solve_all()
    x_gpu = ms.solutions()
    worst = dict(gpu=0.0, oracle=0.0, between=0.0, cond=0.0, bound=0.0)
    for k in range(o.N):
        A = o.matrix(k)
        b = o.rhs(k)
        x_star = spla.spsolve(sp.csc_matrix(A), b)
        x_star = x_star + spla.spsolve(sp.csc_matrix(A), b - A @ x_star)  # one step of iterative refinement
        e_gpu, e_orc = rel_l2(x_gpu[k], x_star), rel_l2(o.solution(k), x_star)
        sv = np.linalg.svd(A, compute_uv=False)
        # radius of the ball the stopping rule allows: |x - x*| <= |A^-1| |g| <= 1e-12 / sigma_min
        radius = 1e-12 / sv[-1] / np.linalg.norm(x_star)
        direct_rounding = 50 * np.finfo(float).eps * sv[0] / sv[-1]
        assert e_gpu <= 2.0 * e_orc + direct_rounding + 1e-13, (
            f"{prm} grid {k}: GPU {e_gpu:.2e} vs oracle {e_orc:.2e} from the direct solution (cond {sv[0] / sv[-1]:.1e})")
        assert e_gpu <= radius + direct_rounding and e_orc <= radius + direct_rounding, (k, e_gpu, e_orc, radius)
        worst["gpu"] = max(worst["gpu"], e_gpu)
        worst["oracle"] = max(worst["oracle"], e_orc)
        worst["between"] = max(worst["between"], rel_l2(x_gpu[k], o.solution(k)))
        worst["cond"] = max(worst["cond"], sv[0] / sv[-1])
        worst["bound"] = max(worst["bound"], radius)
    print(f"\n{prm} {macro_cells} {micro_cells}: distance to the direct solution  GPU {worst['gpu']:.2e}  oracle {worst['oracle']:.2e}  "
          f"GPU-oracle {worst['between']:.2e}  allowed by the stopping rule {worst['bound']:.2e}  cond(A) {worst['cond']:.1e}")
    # The two iterates sit in the same ball around the direct solution, so they differ by at most the sum of their
    # distances to it: THAT is the tolerance of tests/test_gpu_bio.py for these sets (measured on a B200, round 2:
    # GPU-oracle 1.9e-10 / 8.8e-10 at 8 / 16 cells on the curved map against 1e-12 / sigma_min = 1.4e-10 / 3.2e-10).
    assert worst["between"] <= worst["gpu"] + worst["oracle"] + 1e-15
    if micro_cells == 8:
        assert worst["gpu"] + worst["oracle"] < 1e-9  # the bound test_gpu_bio.py uses at this size
    ms.close()
    o.close()
