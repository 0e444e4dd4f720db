#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the CPU oracle (oracle/).

The reference holds no golden vectors for the micro path beyond the seven known answers of
tests/test_tools.cpp (those live in tests/test_reference_known_answers.py).  These fixtures freeze
what the oracle produces — after it has been checked against the SURVEY.md §9 anchors
(tests/test_oracle_anchors.py) — so that (i) a later change of the oracle is noticed and (ii) the GPU
tier can compare against committed numbers.  Run from the repository root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import OracleBio, OracleElliptic, OracleParabolic  # noqa: E402


def elliptic(prm, refs, limit):
    o = OracleElliptic(prm, *refs)
    n = o.run(limit)
    hist = [o.history(c) for c in range(n)]
    out = dict(cycles=n, grid_locations=o.grid_locations(), dof_points=o.dof_points(),
               errors=np.array([h["errors"] for h in hist]), macro=np.array([h["macro"] for h in hist]),
               contribution=np.array([h["contribution"] for h in hist]), micro=np.array([h["micro"] for h in hist]),
               micro_iters=np.array([h["micro_iters"] for h in hist]), table=np.array(o.table()))
    o.close()
    return out


def bio(prm, cells, limit):
    o = OracleBio(prm, *cells, threads=4)
    n = o.run(limit)
    hist = [o.history(c) for c in range(n)]
    out = dict(cycles=n, grid_locations=o.grid_locations(), dof_points=o.dof_points(),
               scalars=np.array([[h["micro_l2"], h["micro_h1"], h["macro_l2"], h["macro_h1"], h["macro_residual"],
                                  h["micro_residual"]] for h in hist]),
               sol_u=np.array([h["sol_u"] for h in hist]), sol_w=np.array([h["sol_w"] for h in hist]),
               cu=np.array([h["cu"] for h in hist]), cw=np.array([h["cw"] for h in hist]),
               micro=np.array([h["micro"] for h in hist]), micro_iters=np.array([h["micro_iters"] for h in hist]))
    o.close()
    return out


def parabolic(prm, level):
    h_inv, t_inv = round(8 * 2 ** (level / 2.0)), round(4 * 2 ** level)
    o = OracleParabolic(prm, h_inv, h_inv, t_inv)
    n = o.run()
    hist = [o.history(s) for s in range(n)]
    out = dict(steps=n, h_inv=h_inv, t_inv=t_inv, passes=o.first_step_passes(), grid_locations=o.grid_locations(),
               scalars=np.array([[h["time"], h["micro_l2"], h["macro_l2"], h["micro_h1"], h["macro_h1"]] for h in hist]),
               pi=np.array([h["pi"] for h in hist]), mass=np.array([h["mass"] for h in hist]),
               rho=np.array([h["rho"] for h in hist]), micro_iters=np.array([h["micro_iters"] for h in hist]))
    o.close()
    return out


CASES = {
    "elliptic_linear_2_3": lambda: elliptic("elliptic_linear.prm", (2, 3), -1),
    "elliptic_non_linear_2_3": lambda: elliptic("elliptic_non_linear.prm", (2, 3), -1),
    "rot_map_test_1_2": lambda: elliptic("rot_map_test.prm", (1, 2), -1),
    "map_test_1_2_k3": lambda: elliptic("map_test.prm", (1, 2), 3),
    "bio_linear_2_4_k4": lambda: bio("linear.prm", (2, 4), 4),
    "bio_biocase_curved_2_4_k4": lambda: bio("biocase-curved.prm", (2, 4), 4),
    "bio_biocase_2_8_k3": lambda: bio("biocase.prm", (2, 8), 3),
    "parabolic_linear_time_0": lambda: parabolic("linear_time.prm", 0),
    "parabolic_full_nonlinear_0": lambda: parabolic("full_nonlinear.prm", 0),
}

if __name__ == "__main__":
    for name, make in CASES.items():
        data = make()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: getattr(v, "shape", v) for k, v in data.items() if k in ("cycles", "steps", "micro")})
