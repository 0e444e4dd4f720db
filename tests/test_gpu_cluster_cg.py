"""GPU tier: micro systems larger than one SM (m > 67) — the thread-block-cluster CG with
distributed-shared-memory halos — against the CPU restatement."""
import os

import numpy as np
import pytest

from helpers import INPUT, OracleBio, OracleElliptic, assert_iterations_consistent, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("micro_cells,clusters", [(70, 2), (96, 3), (128, 4)])
def test_bio_systems_spanning_several_sms(micro_cells, clusters):
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    o = OracleBio("biocase-b.prm", 1, micro_cells)
    xy = o.grid_locations()
    u, w = 1.0 + xy[:, 0], 0.5 - xy[:, 1]
    o.micro_run(u, w)
    ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, "biocase-b.prm")), micro_cells)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_macro_solution(u, w)
    its = ms.run()
    assert ms.stats()["cg_variant"] == 4
    assert (ms.n_dofs + 256 * 17 - 1) // (256 * 17) == clusters
    x = ms.solutions()
    for k in range(o.N):
        assert rel_l2(x[k], o.solution(k)) < 1e-10
    assert_iterations_consistent(o, its)
    cu, cw = ms.coupling()
    cu_o, cw_o = o.flux()
    assert rel_l2(cu, cu_o) < 1e-10 and rel_l2(cw, cw_o) < 1e-10
    ms.close()
    o.close()


def test_elliptic_refine_7():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData

    o = OracleElliptic("elliptic_linear.prm", 1, 7)      # 9 macro DoFs, 128 x 128 micro cells
    xy = o.grid_locations()
    u = 1.0 + 0.5 * xy[:, 0] - 0.25 * xy[:, 1]
    o.micro_run(u)
    ms = MicroSolver(capi.ELLIPTIC, EllipticData(os.path.join(INPUT, "elliptic_linear.prm")), 128)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_macro_solution(u)
    its = ms.run()
    assert ms.stats()["cg_variant"] == 4
    x = ms.solutions()
    for k in range(o.N):
        assert rel_l2(x[k], o.solution(k)) < 1e-10
    assert_iterations_consistent(o, its, unknowns=127 * 127)
    assert rel_l2(ms.coupling(), o.bulk()) < 1e-10
    ms.close()
    o.close()


def test_cluster_and_streaming_kernels_agree():
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import BioData

    g = np.linspace(-0.8, 0.9, 4)
    xy = np.array([(a, b) for b in g for a in g])
    ms = MicroSolver(capi.BIO, BioData(os.path.join(INPUT, "biocase-curved.prm")), 72)
    ms.set_grid_locations(xy)
    ms.setup()
    ms.set_macro_solution(np.cos(xy[:, 0]), np.sin(xy[:, 1]))
    ms.assemble_system()
    out = {}
    for name, env in (("cluster", None), ("global", "global")):
        if env:
            os.environ["MSGPU_CG_VARIANT"] = env
        ms.reload_switches()  # the switches are read when the solver is created
        try:
            ms.zero_solutions()
            its = ms.solve()
            out[name] = (ms.solutions(), its.copy(), ms.stats()["cg_variant"])
        finally:
            os.environ.pop("MSGPU_CG_VARIANT", None)
    assert out["cluster"][2] == 4 and out["global"][2] == 1
    assert np.abs(out["cluster"][0] - out["global"][0]).max() < 1e-9 * np.abs(out["global"][0]).max()
    assert np.abs(out["cluster"][1] - out["global"][1]).max() <= 4
    ms.close()
