"""Committed golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
oracle): the CPU tier checks that the oracle still reproduces them, the GPU tier that the C++ drivers
on the GPU micro path do — cycle by cycle, to 1e-10."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import INPUT, ROOT, OracleBio, OracleElliptic, OracleParabolic, dptr, iptr, rel_l2

GOLD = os.path.join(ROOT, "tests", "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


ELLIPTIC = [("elliptic_linear_2_3", "elliptic_linear.prm", (2, 3), -1),
            ("elliptic_non_linear_2_3", "elliptic_non_linear.prm", (2, 3), -1),
            ("rot_map_test_1_2", "rot_map_test.prm", (1, 2), -1), ("map_test_1_2_k3", "map_test.prm", (1, 2), 3)]
BIO = [("bio_linear_2_4_k4", "linear.prm", (2, 4), 4), ("bio_biocase_curved_2_4_k4", "biocase-curved.prm", (2, 4), 4),
       ("bio_biocase_2_8_k3", "biocase.prm", (2, 8), 3)]
PARABOLIC = [("parabolic_linear_time_0", "linear_time.prm", 0), ("parabolic_full_nonlinear_0", "full_nonlinear.prm", 0)]


# ---------------------------------------------------------------- CPU tier: oracle vs fixtures
@pytest.mark.parametrize("name,prm,refs,limit", ELLIPTIC)
def test_oracle_elliptic_reproduces_fixture(name, prm, refs, limit):
    g = gold(name)
    o = OracleElliptic(prm, *refs)
    assert o.run(limit) == int(g["cycles"])
    for c in range(int(g["cycles"])):
        h = o.history(c)
        assert rel_l2(h["macro"], g["macro"][c]) < 1e-12 and rel_l2(h["micro"], g["micro"][c]) < 1e-12
        np.testing.assert_allclose(h["errors"][[0, 2]], g["errors"][c][[0, 2]], rtol=1e-10)
    assert o.table() == str(g["table"])
    o.close()


@pytest.mark.parametrize("name,prm,cells,limit", BIO)
def test_oracle_bio_reproduces_fixture(name, prm, cells, limit):
    g = gold(name)
    o = OracleBio(prm, *cells, threads=4)
    assert o.run(limit) == int(g["cycles"])
    for c in range(int(g["cycles"])):
        h = o.history(c)
        assert rel_l2(h["sol_u"], g["sol_u"][c]) < 1e-11 and rel_l2(h["micro"], g["micro"][c]) < 1e-11
    o.close()


@pytest.mark.parametrize("name,prm,level", PARABOLIC)
def test_oracle_parabolic_reproduces_fixture(name, prm, level):
    g = gold(name)
    o = OracleParabolic(prm, int(g["h_inv"]), int(g["h_inv"]), int(g["t_inv"]))
    assert o.run() == int(g["steps"]) and o.first_step_passes() == int(g["passes"])
    for s in range(int(g["steps"])):
        h = o.history(s)
        assert rel_l2(h["pi"], g["pi"][s]) < 1e-12 and rel_l2(h["rho"], g["rho"][s]) < 1e-12
    o.close()


# ---------------------------------------------------------------- GPU tier: drivers vs fixtures
def _drivers():
    import torch  # noqa: F401
    from dealiiscale_b200 import capi

    capi.load()
    lib = C.CDLL(os.path.join(ROOT, "dealiiscale_b200", "_build", "libmsdrivers.so"))
    for f in ("msdrv_elliptic_create", "msdrv_bio_create", "msdrv_parabolic_create"):
        getattr(lib, f).restype = C.c_void_p
    lib.msdrv_last_error.restype = C.c_char_p
    return lib


@pytest.mark.gpu
@pytest.mark.parametrize("name,prm,refs,limit", ELLIPTIC)
def test_gpu_elliptic_matches_fixture(name, prm, refs, limit):
    g = gold(name)
    lib = _drivers()
    h = C.c_void_p(lib.msdrv_elliptic_create(os.path.join(INPUT, prm).encode(), *refs))
    assert h.value, lib.msdrv_last_error()
    assert lib.msdrv_elliptic_run(h, limit) == int(g["cycles"])
    N, n = g["micro"].shape[1:]
    for c in range(int(g["cycles"])):
        errs, macro, contrib = np.zeros(4), np.zeros(N), np.zeros(N)
        micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
        lib.msdrv_elliptic_history(h, c, dptr(errs), dptr(macro), dptr(contrib), dptr(micro), iptr(its))
        assert rel_l2(macro, g["macro"][c]) < 1e-10 and rel_l2(micro, g["micro"][c]) < 1e-10
        assert rel_l2(contrib, g["contribution"][c]) < 1e-10
        np.testing.assert_allclose(errs[[0, 2]], g["errors"][c][[0, 2]], rtol=1e-8)
    if limit < 0:
        buf = C.create_string_buffer(1 << 16)
        lib.msdrv_elliptic_table(h, buf, len(buf))
        assert [l.split() for l in buf.value.decode().strip().splitlines()] == \
               [l.split() for l in str(g["table"]).strip().splitlines()]
    lib.msdrv_elliptic_destroy(h)


@pytest.mark.gpu
@pytest.mark.parametrize("name,prm,cells,limit", BIO)
def test_gpu_bio_matches_fixture(name, prm, cells, limit):
    g = gold(name)
    lib = _drivers()
    h = C.c_void_p(lib.msdrv_bio_create(os.path.join(INPUT, prm).encode(), *cells))
    assert h.value, lib.msdrv_last_error()
    assert lib.msdrv_bio_run(h, limit) == int(g["cycles"])
    N, n = g["micro"].shape[1:]
    tol = 1e-8 if "curved" in name else 1e-9
    for c in range(int(g["cycles"])):
        sc = np.zeros(6)
        su, sw, cu, cw = (np.zeros(N) for _ in range(4))
        micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
        lib.msdrv_bio_history(h, c, dptr(sc), dptr(su), dptr(sw), dptr(cu), dptr(cw), dptr(micro), iptr(its))
        assert rel_l2(su, g["sol_u"][c]) < tol and rel_l2(sw, g["sol_w"][c]) < tol
        assert rel_l2(micro, g["micro"][c]) < tol and rel_l2(cu, g["cu"][c]) < tol and rel_l2(cw, g["cw"][c]) < tol
    lib.msdrv_bio_destroy(h)


@pytest.mark.gpu
@pytest.mark.parametrize("name,prm,level", PARABOLIC)
def test_gpu_parabolic_matches_fixture(name, prm, level):
    g = gold(name)
    lib = _drivers()
    h = C.c_void_p(lib.msdrv_parabolic_create(os.path.join(INPUT, prm).encode(), int(g["h_inv"]), int(g["h_inv"]),
                                               int(g["t_inv"])))
    assert h.value, lib.msdrv_last_error()
    assert lib.msdrv_parabolic_run(h, -1) == int(g["steps"])
    assert lib.msdrv_parabolic_first_step_passes(h) == int(g["passes"])
    N, n = g["rho"].shape[1:]
    for s in range(int(g["steps"])):
        sc, pi, mass = np.zeros(5), np.zeros(N), np.zeros(N)
        rho, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
        lib.msdrv_parabolic_history(h, s, dptr(sc), dptr(pi), dptr(mass), dptr(rho), iptr(its))
        assert rel_l2(pi, g["pi"][s]) < 1e-10 and rel_l2(rho, g["rho"][s]) < 1e-10 and rel_l2(mass, g["mass"][s]) < 1e-10
    lib.msdrv_parabolic_destroy(h)
