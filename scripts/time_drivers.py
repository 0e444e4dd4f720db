"""Wall time of the three host drivers at their largest reference sizes (GPU box)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "results"), exist_ok=True)
B = os.path.join(ROOT, "dealiiscale_b200", "_build")
for cmd in ([os.path.join(B, "solve_biomath"), "input/biocase.prm", "6", "6"],
            [os.path.join(B, "solve_biomath"), "input/linear.prm", "5", "6"],
            [os.path.join(B, "solve_elliptic"), "input/elliptic_linear.prm"],
            [os.path.join(B, "solve_parabolic"), "input/linear_time.prm"]):
    for env_extra in ({}, {"MSGPU_HOST_MACRO": "1"}):
        t0 = time.time()
        res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, env=dict(os.environ, MSGPU_SOLUTION_FILES="0", **env_extra))  # time the solvers, not 4225 VTK dumps
        dt = time.time() - t0
        cycles = res.stdout.count("Macro residual") + res.stdout.count("Solving for t")
        print(f"{' '.join(os.path.basename(c) for c in cmd)} {env_extra or ''}: rc {res.returncode}  {dt:.2f} s  "
              f"({cycles} cycles / steps)", flush=True)
        if res.returncode:
            print(res.stderr[-500:])
