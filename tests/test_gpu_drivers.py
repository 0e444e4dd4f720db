"""GPU tier: the C++ host drivers (Manager / TimeManager with the reference's loops and stop tests,
micro path on the GPU through the C ABI) against the oracle's drivers, cycle by cycle
(parity after each of the first K fixed-point cycles / time steps, full history
where the run converges)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import INPUT, ROOT, OracleBio, OracleElliptic, OracleParabolic, dptr, iptr, rel_l2

pytestmark = pytest.mark.gpu


def _drivers():
    import torch  # noqa: F401  (load torch's NCCL before libmsgpu, see dealiiscale_b200/capi.py)
    from dealiiscale_b200 import capi

    capi.load()
    lib = C.CDLL(os.path.join(ROOT, "dealiiscale_b200", "_build", "libmsdrivers.so"))
    for f in ("msdrv_elliptic_create", "msdrv_bio_create", "msdrv_parabolic_create"):
        getattr(lib, f).restype = C.c_void_p
    lib.msdrv_last_error.restype = C.c_char_p
    return lib


def _path(prm):
    return os.path.join(INPUT, prm).encode()


@pytest.mark.parametrize("prm,refs,limit", [
    ("elliptic_linear.prm", (2, 3), -1), ("elliptic_non_linear.prm", (2, 3), -1), ("rot_map_test.prm", (1, 2), -1),
    ("elliptic_linear.prm", (3, 4), -1),
    ("map_test.prm", (2, 3), 3), ("arb_map_test.prm", (2, 3), 3), ("lin_map_test.prm", (1, 2), 3),
])
def test_solve_elliptic_history(prm, refs, limit):
    lib = _drivers()
    o = OracleElliptic(prm, *refs)
    n_cycles = o.run(limit)
    h = C.c_void_p(lib.msdrv_elliptic_create(_path(prm), *refs))
    assert h.value, lib.msdrv_last_error()
    got = lib.msdrv_elliptic_run(h, limit)
    assert got == n_cycles, lib.msdrv_last_error()        # identical cycle count
    N, n = C.c_int(), C.c_int()
    lib.msdrv_elliptic_sizes(h, C.byref(N), C.byref(n))
    N, n = N.value, n.value
    for c in range(n_cycles):
        ref = o.history(c)
        errs, macro, contrib = np.zeros(4), np.zeros(N), np.zeros(N)
        micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
        lib.msdrv_elliptic_history(h, c, dptr(errs), dptr(macro), dptr(contrib), dptr(micro), iptr(its))
        assert rel_l2(macro, ref["macro"]) < 1e-10
        assert rel_l2(contrib, ref["contribution"]) < 1e-10
        assert rel_l2(micro, ref["micro"]) < 1e-10
        np.testing.assert_allclose(errs[[0, 2]], ref["errors"][[0, 2]], rtol=1e-8)     # mL2, ML2
        np.testing.assert_allclose(errs[[1, 3]], ref["errors"][[1, 3]], rtol=1e-5)     # mH1, MH1 (finite differences)
    if limit < 0:
        buf = C.create_string_buffer(1 << 16)
        lib.msdrv_elliptic_table(h, buf, len(buf))
        ours = [l.split() for l in buf.value.decode().strip().splitlines()]
        theirs = [l.split() for l in o.table().strip().splitlines()]
        assert ours == theirs                                 # printed convergence table, 3 significant digits
    lib.msdrv_elliptic_destroy(h)
    o.close()


@pytest.mark.parametrize("prm,cells,limit", [
    ("linear.prm", (4, 16), -1), ("biocase-curved.prm", (3, 8), 6), ("biocase.prm", (4, 8), 5),
    ("biocase-a.prm", (2, 8), 8),
])
def test_solve_biomath_history(prm, cells, limit):
    lib = _drivers()
    o = OracleBio(prm, *cells, threads=8)
    n_cycles = o.run(limit)
    h = C.c_void_p(lib.msdrv_bio_create(_path(prm), *cells))
    assert h.value, lib.msdrv_last_error()
    assert lib.msdrv_bio_run(h, limit) == n_cycles, lib.msdrv_last_error()
    N, n = C.c_int(), C.c_int()
    lib.msdrv_bio_sizes(h, C.byref(N), C.byref(n))
    N, n = N.value, n.value
    tol = 1e-8 if prm == "biocase-curved.prm" else 1e-9       # warm-started, ill-conditioned systems: see test_gpu_bio
    for c in range(n_cycles):
        ref = o.history(c)
        sc = np.zeros(6)
        su, sw, cu, cw = (np.zeros(N) for _ in range(4))
        micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
        lib.msdrv_bio_history(h, c, dptr(sc), dptr(su), dptr(sw), dptr(cu), dptr(cw), dptr(micro), iptr(its))
        assert rel_l2(su, ref["sol_u"]) < tol and rel_l2(sw, ref["sol_w"]) < tol
        assert rel_l2(cu, ref["cu"]) < tol and rel_l2(cw, ref["cw"]) < tol
        assert rel_l2(micro, ref["micro"]) < tol
        if o.has_solution():
            np.testing.assert_allclose(sc[[0, 2]], [ref["micro_l2"], ref["macro_l2"]], rtol=1e-7)
        else:
            np.testing.assert_allclose(sc[4:6], [ref["macro_residual"], ref["micro_residual"]], rtol=1e-4, atol=1e-12)
    lib.msdrv_bio_destroy(h)
    o.close()


@pytest.mark.parametrize("prm,level", [("linear_time.prm", 0), ("linear_time.prm", 2), ("full_nonlinear.prm", 1)])
def test_solve_parabolic_history(prm, level):
    lib = _drivers()
    h_inv, t_inv = round(8 * 2 ** (level / 2.0)), round(4 * 2 ** level)
    o = OracleParabolic(prm, h_inv, h_inv, t_inv)
    steps = o.run()
    h = C.c_void_p(lib.msdrv_parabolic_create(_path(prm), h_inv, h_inv, t_inv))
    assert h.value, lib.msdrv_last_error()
    assert lib.msdrv_parabolic_run(h, -1) == steps, lib.msdrv_last_error()
    assert lib.msdrv_parabolic_first_step_passes(h) == o.first_step_passes()
    N, n = C.c_int(), C.c_int()
    lib.msdrv_parabolic_sizes(h, C.byref(N), C.byref(n))
    N, n = N.value, n.value
    for s in range(steps):
        ref = o.history(s)
        sc, pi, mass = np.zeros(5), np.zeros(N), np.zeros(N)
        rho, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
        lib.msdrv_parabolic_history(h, s, dptr(sc), dptr(pi), dptr(mass), dptr(rho), iptr(its))
        assert sc[0] == ref["time"]
        assert rel_l2(pi, ref["pi"]) < 1e-10
        assert rel_l2(rho, ref["rho"]) < 1e-10
        assert rel_l2(mass, ref["mass"]) < 1e-10
        # linear exact solutions are reproduced to rounding (errors ~1e-11): compare those absolutely
        np.testing.assert_allclose(sc[1:3], [ref["micro_l2"], ref["macro_l2"]], rtol=1e-7, atol=1e-9)
    lib.msdrv_parabolic_destroy(h)
    o.close()


def test_cli_binaries_run_and_write_the_table(tmp_path):
    exe = os.path.join(ROOT, "dealiiscale_b200", "_build", "solve_biomath")
    os.makedirs(os.path.join(tmp_path, "results"), exist_ok=True)
    os.symlink(INPUT, os.path.join(tmp_path, "input"))
    r = subprocess.run([exe, "input/linear.prm", "1", "3", "0"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Results: 'macro_refinement', 'micro_refinement', 'num_threads', 'wall_time', 'cpu_time'" in r.stdout
    table = open(os.path.join(tmp_path, "results", "out.txt")).read()
    assert table.split()[:7] == ["cycle", "cells", "dofs", "mL2", "mH1", "ML2", "MH1"]
    # the solution files of bio manager.cpp:125-195 (refinements 1, 3: 2x2 macro cells -> 9 micro grids of 8x8 cells)
    res = os.path.join(tmp_path, "results")
    for f in ["u-solution.vtk", "w-solution.vtk", "micro-error.vtk", "micro-exact.vtk", "micro-solutions/grid_locations.txt"] \
            + [f"micro-solutions/v-solution-{k}.vtk" for k in range(9)]:
        assert os.path.getsize(os.path.join(res, f)) > 0, f
    assert len(open(os.path.join(res, "micro-solutions", "grid_locations.txt")).read().split()) == 18
    from test_output_files import parse_vtk

    pts, cells, types, name, data = parse_vtk(os.path.join(res, "micro-exact.vtk"))
    assert name == "v" and len(cells) == 64 and np.isfinite(data).all()
    # linear.prm has a manufactured solution: the projected exact solution minus the computed one is small
    _, _, _, _, err = parse_vtk(os.path.join(res, "micro-error.vtk"))
    assert np.abs(err).max() < 0.5 * np.abs(data).max()


def test_elliptic_and_parabolic_binaries_write_their_solution_files(tmp_path):
    os.makedirs(os.path.join(tmp_path, "results"), exist_ok=True)
    os.symlink(INPUT, os.path.join(tmp_path, "input"))
    exe = os.path.join(ROOT, "dealiiscale_b200", "_build", "solve_parabolic")
    r = subprocess.run([exe, "input/linear_time.prm", "1"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    res = os.path.join(tmp_path, "results")
    for f in ("final_macro_solution.gpl", "final_micro_solution.gpl", "patched_micro_solution.gpl",
              "micro-solution001.vtk", "macro-solution002.vtk", "linear_time_convergence_table.txt"):
        assert os.path.getsize(os.path.join(res, f)) > 0, f
    # 81 micro grids of 8x8 cells, each with its own gnuplot header (rho_solver.cpp:249-265)
    patched = open(os.path.join(res, "patched_micro_solution.gpl")).read()
    assert patched.count("# <x> <y> <solution> ") == 81
    exe = os.path.join(ROOT, "dealiiscale_b200", "_build", "solve_elliptic")
    r = subprocess.run([exe, "input/elliptic_linear.prm"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    for f in ("macro-solution.gpl", "micro-computed.gpl", "micro-error.gpl", "micro-exact.gpl",
              "elliptic_linear_convergence_table.txt"):
        assert os.path.getsize(os.path.join(res, f)) > 0, f
    rows = [ln.split() for ln in open(os.path.join(res, "micro-error.gpl")) if ln.strip() and ln[0] != "#"]
    err = np.array(rows, dtype=float)[:, 2]
    assert len(err) == 4 * 32 * 32 and np.abs(err).max() < 1e-2  # finest level: computed - projected exact


def test_device_and_host_macro_paths_agree():
    """the drivers keep the macro systems on the device by default
    (msgpu_macro_*); MSGPU_HOST_MACRO=1 solves them on the host as the reference does."""
    lib = _drivers()
    prm, cells, limit = "linear.prm", (4, 8), 4
    out = {}
    for name, env in (("device", None), ("host", "1")):
        if env:
            os.environ["MSGPU_HOST_MACRO"] = env
        try:
            h = C.c_void_p(lib.msdrv_bio_create(_path(prm), *cells))
            assert h.value, lib.msdrv_last_error()
            n_cycles = lib.msdrv_bio_run(h, limit)
            N, n = C.c_int(), C.c_int()
            lib.msdrv_bio_sizes(h, C.byref(N), C.byref(n))
            N, n = N.value, n.value
            sc = np.zeros(6)
            u, w, cu, cw = np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N)
            micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
            lib.msdrv_bio_history(h, n_cycles - 1, dptr(sc), dptr(u), dptr(w), dptr(cu), dptr(cw), dptr(micro), iptr(its))
            out[name] = (n_cycles, u.copy(), w.copy(), micro.copy())
            lib.msdrv_bio_destroy(h)
        finally:
            os.environ.pop("MSGPU_HOST_MACRO", None)
    assert out["device"][0] == out["host"][0]
    for a, b in zip(out["device"][1:], out["host"][1:]):
        assert rel_l2(a, b) < 1e-10


@pytest.mark.parametrize("prm", ["linear_time.prm", "full_nonlinear.prm"])
def test_pi_solver_load_on_the_device_equals_the_host_assembly(prm):
    """PiSolver::assemble_system (pi_solver.cpp:76-122): the load term  int phi_i rho_h pi_h  - a product of two
    finite-element functions, with min(|m|, 1) and theta*min(|pi|, sqrt|pi|) in the nonlinear set - is evaluated by
    k_macro_product on the device from the device-resident micro masses and the previous macro solution;
    MSGPU_HOST_PRODUCT=1 keeps the reference's host loop.  Same histories, step by step."""
    lib = _drivers()
    h_inv, t_inv = 16, 8
    out = {}
    for name, env in (("device", None), ("host", "1")):
        if env:
            os.environ["MSGPU_HOST_PRODUCT"] = env
        try:
            h = C.c_void_p(lib.msdrv_parabolic_create(_path(prm), h_inv, h_inv, t_inv))
            assert h.value, lib.msdrv_last_error()
            steps = lib.msdrv_parabolic_run(h, -1)
            assert steps >= 4, lib.msdrv_last_error()
            N, n = C.c_int(), C.c_int()
            lib.msdrv_parabolic_sizes(h, C.byref(N), C.byref(n))
            N, n = N.value, n.value
            hist = []
            for s in range(steps):
                sc, pi, mass = np.zeros(5), np.zeros(N), np.zeros(N)
                rho, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
                lib.msdrv_parabolic_history(h, s, dptr(sc), dptr(pi), dptr(mass), dptr(rho), iptr(its))
                hist.append((pi.copy(), mass.copy(), rho.copy()))
            out[name] = (lib.msdrv_parabolic_first_step_passes(h), hist)
            lib.msdrv_parabolic_destroy(h)
        finally:
            os.environ.pop("MSGPU_HOST_PRODUCT", None)
    assert out["device"][0] == out["host"][0] and len(out["device"][1]) == len(out["host"][1])
    for (pa, ma, ra), (pb, mb, rb) in zip(out["device"][1], out["host"][1]):
        assert rel_l2(pa, pb) < 1e-11 and rel_l2(ma, mb) < 1e-11 and rel_l2(ra, rb) < 1e-11


@pytest.mark.parametrize("prm,cells,limit", [("linear.prm", (4, 8), 4), ("biocase.prm", (4, 8), 4)])
def test_stop_test_on_the_device_equals_the_host_aggregation(prm, cells, limit):
    """The macro-aggregated error / residual norms that decide the stop (bio manager.cpp:70-109) are computed by
    msgpu_norm_* without copying the per-grid scalars to the host: same cells, same points, same order and no FMA
    contraction, so the scalars of every cycle EQUAL those of the host loop (MSGPU_HOST_STOP_TEST=1) bit for bit."""
    lib = _drivers()
    out = {}
    for name, env in (("device", None), ("host", "1")):
        if env:
            os.environ["MSGPU_HOST_STOP_TEST"] = env
        try:
            h = C.c_void_p(lib.msdrv_bio_create(_path(prm), *cells))
            assert h.value, lib.msdrv_last_error()
            n_cycles = lib.msdrv_bio_run(h, limit)
            N, n = C.c_int(), C.c_int()
            lib.msdrv_bio_sizes(h, C.byref(N), C.byref(n))
            N, n = N.value, n.value
            scalars = []
            for c in range(n_cycles):
                sc = np.zeros(6)
                u, w, cu, cw = np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N)
                micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
                lib.msdrv_bio_history(h, c, dptr(sc), dptr(u), dptr(w), dptr(cu), dptr(cw), dptr(micro), iptr(its))
                scalars.append(sc.copy())
            out[name] = (n_cycles, np.array(scalars))
            lib.msdrv_bio_destroy(h)
        finally:
            os.environ.pop("MSGPU_HOST_STOP_TEST", None)
    assert out["device"][0] == out["host"][0]
    assert np.array_equal(out["device"][1], out["host"][1]), (out["device"][1], out["host"][1])
    assert np.abs(out["device"][1]).max() > 0


@pytest.mark.parametrize("prm,cells,limit", [("biocase.prm", (4, 8), 5), ("biocase-curved.prm", (3, 6), 4)])
def test_assembly_beside_the_macro_solve_changes_nothing(prm, cells, limit):
    """msgpu_assemble_begin puts the macro-independent part of the micro assembly (moments, gather) on a second stream
    while the macro systems are solved; the reference does one after the other (bio manager.cpp:65-68).  Same kernels
    on the same inputs: every cycle's scalars, macro solutions, flux integrals, micro solutions and CG step counts must
    EQUAL those of the sequential order (MSGPU_OVERLAP_ASSEMBLY=0)."""
    lib = _drivers()
    out = {}
    for name, env in (("beside", None), ("sequential", "0")):
        if env:
            os.environ["MSGPU_OVERLAP_ASSEMBLY"] = env
        try:
            h = C.c_void_p(lib.msdrv_bio_create(_path(prm), *cells))
            assert h.value, lib.msdrv_last_error()
            n_cycles = lib.msdrv_bio_run(h, limit)
            N, n = C.c_int(), C.c_int()
            lib.msdrv_bio_sizes(h, C.byref(N), C.byref(n))
            N, n = N.value, n.value
            rec = []
            for c in range(n_cycles):
                sc = np.zeros(6)
                u, w, cu, cw = np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N)
                micro, its = np.zeros((N, n)), np.zeros(N, dtype=np.int32)
                lib.msdrv_bio_history(h, c, dptr(sc), dptr(u), dptr(w), dptr(cu), dptr(cw), dptr(micro), iptr(its))
                rec.append((sc.copy(), u, w, cu, cw, micro, its))
            out[name] = rec
            lib.msdrv_bio_destroy(h)
        finally:
            os.environ.pop("MSGPU_OVERLAP_ASSEMBLY", None)
    assert len(out["beside"]) == len(out["sequential"]) >= 2
    for a, b in zip(out["beside"], out["sequential"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert np.abs(out["beside"][-1][5]).max() > 0
