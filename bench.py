#!/usr/bin/env python
"""Benchmark of the batched micro path (BASELINE.json metric: micro DoF-solves/s per fixed-point
iteration; % of HBM roofline).

One "step" = one fixed-point iteration of the micro path of `solve_biomath <prm> 6 6`
(reference: src/bio-multiscale/micro.cpp:292-305 + macro.cpp:299-306): for every macro DoF owned
by the rank, assemble the micro system (64x64 Q1 cells, QGauss(12)), solve it with CG from a zero
start vector (abs. tol 1e-12, as the first cycle of the reference does), and compute the two
micro->macro flux integrals; with N>1 ranks the integrals are allgathered and the macro solution
is broadcast (NCCL).  Weak scaling: every GPU owns 65x65 = 4225 macro points.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

`--impl reference` times the CPU restatement of the reference algorithm (oracle/; deal.II is not
installable here) on the host cores for the same metric on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "micro_dof_solves_per_s"
UNIT = "DoF-solves/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--prm", default="biocase.prm")
    ap.add_argument("--macro-ref", type=int, default=6, help="macro refinement: 2^r cells per axis per GPU")
    ap.add_argument("--micro-ref", type=int, default=6, help="micro refinement: 2^r cells per axis")
    ap.add_argument("--cpu-sample", type=int, default=0, help="micro systems in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precond", default="identity", choices=["identity", "jacobi", "ssor"],
                    help="A/B only: the reference solves with PreconditionIdentity, which is what the headline number uses")
    ap.add_argument("--nccl-allgather", action="store_true",
                    help="multi-GPU: exchange the flux integrals with ncclAllGather instead of the fused peer-memory kernel")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the headline): 4225 macro points per GPU; strong: the 4225 macro points of the "
                         "literal `6 6` lattice shared by all GPUs (also reported as the 'strong' sub-record of a weak run)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the 'other_configs' records (parabolic step, elliptic S2 workloads; 1 GPU only)")
    ap.add_argument("--no-cycle", action="store_true",
                    help="skip the 'cycle' / 'strong' records (the C++ driver solve_biomath on the literal lattice)")
    ap.add_argument("--cycles", type=int, default=6, help="fixed-point cycles the C++ driver runs for the 'cycle' record")
    return ap.parse_args()


def macro_points(macro_cells: int, n_gpus: int, strong: bool = False):
    """Macro support points: a (macro_cells+1) x (macro_cells+1) block per GPU, side by side in x0,
    the whole lattice scaled into [-1,1]^2.  Row-major blocks so that every rank owns a contiguous
    range of macro DoFs.  strong: ONE such block for the whole job (the literal lattice of the reference)."""
    if strong:
        n_gpus = 1
    m = macro_cells + 1
    x1 = np.linspace(-1.0, 1.0, m)
    x0 = np.linspace(-1.0, 1.0, m * n_gpus)
    pts = []
    for g in range(n_gpus):
        for j in range(m):
            for i in range(m):
                pts.append((x0[g * m + i], x1[j]))
    return np.array(pts)


def macro_values(xy):
    """Synthetic macro iterate (bounded, smooth): the shape of the exact macro solutions of linear.prm."""
    u = xy[:, 1] * np.sin(xy[:, 0])
    w = xy[:, 0] * np.cos(xy[:, 1])
    return np.ascontiguousarray(u), np.ascontiguousarray(w)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe), through
    NVML (same counters as `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`), every 10 ms."""

    REASONS = {  # nvmlClocksEventReason* bit masks
        "hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
        "hw_power_brake_slowdown": 0x80,
    }

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._handle = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
        for name, bit in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        parts = [p.strip() for p in out.split(",")]
        if len(parts) >= 6:
            self.samples.append(float(parts[0]))
            self.max_mhz = float(parts[1])
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    self.reasons.add(nm)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._handle is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._handle is not None else 0.2)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "source": "nvml" if self._handle is not None else "nvidia-smi"}


CG_KERNEL_NAMES = {
    0: "k_cg_resident (batched CG, one CTA per micro system, matrix resident in shared memory)",
    1: "k_cg_global (batched CG, matrix streamed from L2/HBM)",
    2: "k_cg_strip (batched CG, one CTA per micro system, matrix resident in shared memory)",
    3: "k_cg_strip_tmem (batched CG, one CTA per micro system, matrix resident in tensor memory + shared memory)",
    4: "k_cg_cluster (batched CG, one thread-block cluster per micro system, matrix resident in tensor memory)",
    5: "k_cg_stencil (batched CG for class-uniform systems: couplings in registers, x/g/h in tensor memory, "
       "two micro systems per SM)",
}


def fp64_record(cg_flops, cg_ms, n_sms, sm_mhz):
    """Achieved FP64 rate of the CG kernel against what the FP64 pipe can issue (64 FMA lanes per SM and clock)."""
    peak = n_sms * 64 * 2 * sm_mhz * 1e6 / 1e12
    ach = cg_flops / (cg_ms * 1e-3) / 1e12 if cg_ms > 0 else 0.0
    return {"achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": f"{n_sms} SMs x 64 FP64 FMA/clk x {sm_mhz:.0f} MHz (nominal; ncu sm__inst_executed_pipe_fp64)",
            "flops_model": "sum over systems of CG steps x (2 nnz + 10 n): SpMV, two dot products, three vector updates"}


def algorithmic_bytes(n, nnz, iters, n_systems_total):
    """SURVEY.md §8d byte model.  Returns (whole step, CG kernel only)."""
    its = np.asarray(iters, dtype=np.float64)
    step = (8.0 * nnz * (1.0 + its) + 24.0 * n + 4.0 * (nnz + n + 1) / max(1, n_systems_total)).sum()
    cg = (8.0 * nnz * its + 24.0 * n).sum()
    return float(step), float(cg)


def cpu_baseline(prm_path, macro_cells, micro_cells, sample, threads):
    """The CPU restatement (oracle) timed on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import OracleBio  # test infrastructure: the checker / baseline, never the product

    # a macro mesh just large enough to hold `sample` grids keeps the oracle's memory small
    cells = 1
    while (cells + 1) ** 2 < sample:
        cells += 1
    o = OracleBio(prm_path, cells, micro_cells, threads)
    xy = o.grid_locations()
    u, w = macro_values(xy)
    sample = min(sample, o.N)
    secs = o.time_micro(u, w, sample, threads)
    its = o.iters()[:sample]
    n = o.n
    o.close()
    return {"value": sample * n / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample} of the {micro_cells}x{micro_cells}-cell micro systems (assemble + CG + flux integrals), "
                      f"{threads} threads over systems, {secs:.2f} s; CPU restatement of the reference algorithm "
                      f"(deal.II unavailable)",
            "seconds": secs, "mean_cg_iterations": float(np.mean(its))}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    micro_cells = 2 ** args.micro_ref
    prm_path = os.path.join(ROOT, "input", args.prm)
    # bounded sample per step: the whole --steps/--warmup run should end within ~3 minutes
    probe = cpu_baseline(prm_path, 2 ** args.macro_ref, micro_cells, 8 * threads, threads)
    systems_per_s = 8 * threads / max(probe["seconds"], 1e-3)
    budget = 150.0 / (args.steps + 0.5 * args.warmup + 1e-9)
    sample = args.cpu_sample or int(min(128 * threads, max(threads, systems_per_s * budget)))
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(prm_path, 2 ** args.macro_ref, micro_cells, max(1, sample // 2), threads)
    t0 = time.time()
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(prm_path, 2 ** args.macro_ref, micro_cells, sample, threads)
        vals.append(last["value"])
    total = time.time() - t0
    v = float(np.mean(vals))
    last["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    mc, Mc = 2 ** args.micro_ref, 2 ** args.macro_ref
    return {
        "workload": f"solve_biomath {args.prm} {args.macro_ref} {args.micro_ref}: one fixed-point iteration of the micro "
                    f"path (assemble + CG from zero start + flux integrals)",
        "prm": args.prm, "micro_cells_per_axis": mc, "micro_dofs": (mc + 1) ** 2, "q_degree": 12,
        "macro_points_per_gpu": (Mc + 1) ** 2 if args.scaling == "weak" else -(-(Mc + 1) ** 2 // n_gpus),
        "macro_points_total": (Mc + 1) ** 2 * (n_gpus if args.scaling == "weak" else 1),
        "cg": f"deal.II recurrence, {args.precond} preconditioner, abs tol 1e-12, max 10000",
        "parallelism": (f"macro points sharded over {n_gpus} GPU(s); flux integrals exchanged "
                        + ("with ncclAllGather" if args.nccl_allgather else
                           "by the coupling kernel itself (peer-memory stores over NVLink + flag per peer)")
                        + "; e2e adds an ncclBroadcast of the macro iterate per step") if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2: 0.7 GB of matrix values + 0.55 GB of cell load moments written and read per step",
    }


def other_configs(args, torch, device, clocks):
    """The BASELINE.json configs the headline does not run, at their S2 size (4225 grids x 4225 DoFs) where they have
    one: one pass of the micro path each (what `step` means for that driver), device-resident, CUDA events on the
    launching stream; CPU restatement on a bounded sample beside it."""
    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver
    from dealiiscale_b200.prm import EllipticData, TwoPressureData

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    from helpers import OracleElliptic, OracleParabolic, dptr  # the checker / baseline, never the product

    g = np.linspace(-1.0, 1.0, 65)
    X, Y = np.meshgrid(g, g, indexing="ij")
    xy = np.ascontiguousarray(np.stack([X.ravel(), Y.ravel()], axis=1))
    n_sms = torch.cuda.get_device_properties(device).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    steps = max(2, min(args.steps, 5))
    out = []
    specs = [
        ("parabolic", "linear_time.prm", "solve_parabolic linear_time.prm, finest level (h_inv 64, dt 1/128): one implicit-Euler "
                                         "micro step of 4225 grids x 4225 DoFs (rho_solver.cpp:232-237) + get_micro_mass"),
        ("parabolic", "full_nonlinear.prm", "solve_parabolic full_nonlinear.prm, finest level (h_inv 64, dt 1/128): one micro step "
                                            "+ get_micro_mass"),
        ("elliptic", "elliptic_non_linear.prm", "solve_elliptic elliptic_non_linear.prm, S2 (refinement 6 / 6): y-dependent Jacobian, "
                                                "4225 DISTINCT 4225-DoF systems, general assembly + general CG kernel"),
        ("elliptic", "map_test.prm", "solve_elliptic map_test.prm, S2 (refinement 6 / 6): per-point linear maps, one stiffness "
                                     "stencil per system, Dirichlet data differ"),
    ]
    for kind, prm, label in specs:
        path = os.path.join(ROOT, "input", prm)
        rec = {"config": label, "prm": prm, "grids": int(xy.shape[0])}
        try:
            if kind == "parabolic":
                ms = MicroSolver(capi.PARABOLIC, TwoPressureData(path), 64, device=device)
            else:
                ms = MicroSolver(capi.ELLIPTIC, EllipticData(path), 64, device=device)
            ms.use_stream(torch.cuda.current_stream().cuda_stream)
            ms.set_grid_locations(xy)
            ms.setup()
            n = ms.n_dofs
            nnz = (3 * 65 - 2) ** 2
            u = np.exp(0.25 * (xy[:, 0] ** 2 + xy[:, 1] ** 2)) if kind == "parabolic" else np.sin(xy[:, 0]) * xy[:, 1] + 0.3
            ms.set_macro_solution(u)
            dt, t = 1.0 / 128, 0.0
            no_conv = 0

            def one():
                nonlocal t, no_conv
                try:
                    if kind == "parabolic":
                        t += dt
                        its_ = ms.iterate(dt, t)
                    else:
                        its_ = ms.run()
                except capi.MsgpuError as err:  # SolverControl's 1000 steps: reached in the reference as well
                    if err.code != capi.ERR_NO_CONVERGENCE:
                        raise
                    no_conv += 1
                    its_ = ms.last_iterations
                ms.coupling_on_device()
                return its_

            for _ in range(2):
                one()
            ms.enable_timing(True)
            ms.stats(reset=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                its = one()
            e1.record()
            torch.cuda.synchronize()
            ms_step = e0.elapsed_time(e1) / steps
            st = ms.stats()
            cg_ms = st["ms_cg"] / steps
            cg_flops = float(its.sum()) * (2.0 * nnz + 10.0 * n)
            rec.update({
                "metric": METRIC, "value": xy.shape[0] * n / (ms_step * 1e-3), "unit": UNIT, "ms_per_step": ms_step,
                "steps": steps, "phases_ms_per_step": {k: st[k] / steps for k in
                                                       ("ms_tables", "ms_moments", "ms_gather", "ms_cg", "ms_coupling")},
                "cg_iterations": {"mean": float(its.mean()), "max": int(its.max()), "min": int(its.min())},
                "cg_kernel": CG_KERNEL_NAMES.get(int(st["cg_variant"]), "batched CG"),
                "fp64": fp64_record(cg_flops, cg_ms, n_sms, sm_mhz),
                "solver_control_limit_hit_in_steps": no_conv,
            })
            ms.close()
            torch.cuda.synchronize()
            if not args.no_cpu_baseline:
                t0 = time.time()
                if kind == "parabolic":
                    o = OracleParabolic(path, 2, 64, 128)  # 9 grids of 64 x 64 cells, one thread as solve_parabolic runs
                    pi = np.exp(0.25 * (o.grid_locations() ** 2).sum(axis=1))
                    t1 = time.time()
                    o.micro_step(pi, dt, dt)
                    secs = time.time() - t1
                    rec["cpu_baseline"] = {"value": o.N * o.n / secs, "unit": UNIT, "cores": 1, "kind": "port",
                                           "sample": f"{o.N} of the 4225 grids, one micro step, 1 thread, {secs:.2f} s; CPU "
                                                     "restatement of the reference algorithm (deal.II unavailable)"}
                    o.close()
                else:
                    o = OracleElliptic(path, 1, 6)  # 9 systems of 64 x 64 cells, one thread as solve_elliptic runs
                    o.lib.orc_ell_time_micro.restype = C.c_double
                    uu = np.sin(o.grid_locations()[:, 0]) * o.grid_locations()[:, 1] + 0.3
                    k1 = 3
                    secs = o.lib.orc_ell_time_micro(o.h, dptr(uu), 0, k1, 1)
                    rec["cpu_baseline"] = {"value": k1 * o.n / secs, "unit": UNIT, "cores": 1, "kind": "port",
                                           "sample": f"{k1} of the 4225 systems (assemble + CG + bulk integral), 1 thread, "
                                                     f"{secs:.2f} s; dense pull-back table; CPU restatement of the reference "
                                                     "algorithm (deal.II unavailable)"}
                    o.close()
                    if prm == "map_test.prm":
                        # what the original pays: the same path with the string-keyed MapMap in the inner loops
                        # (src/tools/mapping.cpp:15-57), at 32 x 32 cells to stay within seconds
                        o5 = OracleElliptic(path, 1, 5)
                        o5.lib.orc_ell_time_micro.restype = C.c_double
                        o5.lib.orc_ell_time_micro_mapmap.restype = C.c_double
                        u5 = np.sin(o5.grid_locations()[:, 0]) * o5.grid_locations()[:, 1] + 0.3
                        dense = o5.lib.orc_ell_time_micro(o5.h, dptr(u5), 0, 2, 1)
                        build = C.c_double()
                        strm = o5.lib.orc_ell_time_micro_mapmap(o5.h, dptr(u5), 0, 2, C.byref(build))
                        rec["cpu_baseline"]["with_mapmap"] = {
                            "micro_cells_per_axis": 32, "systems": 2, "dense_table_s": dense, "string_map_s": strm,
                            "string_map_build_s": build.value, "slowdown": strm / dense,
                            "value": 2 * o5.n / strm, "unit": UNIT,
                            "note": "std::map<std::string, ...> lookups with ostringstream keys per quadrature point, as in "
                                    "MicroSolver::integrate_cell / MacroSolver::get_micro_bulk of the reference"}
                        o5.close()
                rec["cpu_baseline"]["wall_s_including_setup"] = time.time() - t0
        except Exception as err:  # a config that cannot run must not take the headline down with it
            rec["error"] = f"{type(err).__name__}: {err}"
        out.append(rec)
    return out


def driver_cycle(args, rank, world, dist, reuse_assembly=False):
    """`solve_biomath <prm> 6 6` - the C++ host driver on top of the C ABI, the whole fixed-point loop of the reference
    (src/bio-multiscale/manager.cpp:51-80): macro assemble + solve, micro assemble + solve, coupling integrals with
    their exchange, stop test - on the LITERAL lattice (4225 macro points in total), one process per GPU started by
    every rank of this job.  Strong scaling of the real thing; per-cycle times from the driver's own clock."""
    import re
    import shutil
    import tempfile

    exe = os.path.join(ROOT, "dealiiscale_b200", "_build", "solve_biomath")
    if not os.path.exists(exe):
        return {"cycle": {"error": "solve_biomath is not built"}, "strong": None} if rank == 0 else None
    box = [tempfile.mkdtemp(prefix="msgpu-bench-") if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(box, src=0)
    env = dict(os.environ)
    # the 4225 VTK files of the micro solutions (0.85 GB, ~35 s of the driver's wall time) are not part of the loop
    env.update({"MSGPU_DRIVER_TIMING": "1", "MSGPU_CYCLE_LIMIT": str(args.cycles), "MSGPU_RENDEZVOUS_DIR": box[0],
                "MSGPU_SOLUTION_FILES": "0", "MSGPU_REUSE_ASSEMBLY": "1" if reuse_assembly else "0"})
    cwd = os.path.join(box[0], f"rank{rank}")
    os.makedirs(cwd, exist_ok=True)
    t0 = time.time()
    try:
        res = subprocess.run([exe, os.path.join(ROOT, "input", args.prm), str(args.macro_ref), str(args.micro_ref)],
                             cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
        rc, text = res.returncode, res.stdout
        err_text = res.stderr[-400:]
    except subprocess.TimeoutExpired:
        rc, text, err_text = -1, "", "timeout"
    wall = time.time() - t0
    if dist is not None:
        dist.barrier()
    out = None
    if rank == 0:
        macro = [float(v) for v in re.findall(r"\[timing\] macro ([0-9.]+) ms", text)]
        micro = [float(v) for v in re.findall(r"micro ([0-9.]+) ms", text)]
        steps = [float(v) for v in re.findall(r"mean CG steps ([0-9.]+)", text)]
        stop = [float(v) for v in re.findall(r"\[timing\] stop test ([0-9.]+) ms", text)]
        phases = re.search(r"\[timing\] construct ([0-9.]+) s, setup ([0-9.]+) s, run \+ output ([0-9.]+) s", text)
        k = min(len(macro), len(micro), len(stop))
        if rc != 0 or k < 2:
            out = {"cycle": {"error": f"solve_biomath rc {rc}: {err_text}"}, "strong": None}
        else:
            cyc = [macro[i] + micro[i] + stop[i] for i in range(k)]
            # cycle 1 starts the micro CG from zero (what the weak-scaling step measures); later cycles warm-start
            Mc, mc = 2 ** args.macro_ref, 2 ** args.micro_ref
            work = float((Mc + 1) ** 2) * float((mc + 1) ** 2)
            later = slice(1, k)
            out = {
                "cycle": {
                    "workload": f"solve_biomath {args.prm} {args.macro_ref} {args.micro_ref} (C++ driver, {world} process(es), one per GPU), "
                                f"{k} fixed-point cycles: macro assemble+solve, micro assemble+solve+flux integrals+exchange, stop test",
                    "ms_macro": macro[:k], "ms_micro": micro[:k], "ms_stop_test": stop[:k], "ms_cycle": cyc,
                    "mean_cg_steps": steps[:k],
                    "first_cycle_ms": cyc[0], "later_cycles_ms_median": float(np.median(cyc[later])),
                    "driver_wall_s": wall,
                    "driver_phases_s": {"construct": float(phases.group(1)), "setup": float(phases.group(2)),
                                        "run_and_output": float(phases.group(3))} if phases else None,
                },
                "strong": {
                    "scaling": "strong", "n_gpus": world, "macro_points_total": (Mc + 1) ** 2,
                    "metric": METRIC, "unit": UNIT,
                    # whole cycles after the first (the first also pays one-off allocations and lazy module loads)
                    "value": work / (float(np.median(cyc[later])) * 1e-3),
                    "ms_cycle": float(np.median(cyc[later])),
                    "ms_micro": float(np.median(micro[later])), "ms_macro": float(np.median(macro[later])),
                    "ms_stop_test": float(np.median(stop[later])),
                    "micro_phase_value": work / (float(np.median(micro[later])) * 1e-3),
                    "value_later_cycles": work / (float(np.median(cyc[later])) * 1e-3),
                    "value_first_cycle": work / (cyc[0] * 1e-3),
                    "micro_phase_value_first_cycle": work / (micro[0] * 1e-3),
                    "ms_cycle_first": cyc[0], "ms_micro_first": micro[0],
                    "note": "whole fixed-point cycle of the real driver on the literal lattice (4225 macro points in total, "
                            "sharded over the GPUs of the job); medians over the cycles after the first, which warm-start "
                            "(fewer CG steps than the zero start of cycle 1 and of the weak-scaling step)",
                },
            }
        shutil.rmtree(box[0], ignore_errors=True)
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the micro path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    # NCCL prints its version banner (and any debug output) on stdout, ahead of the JSON line: send it to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from dealiiscale_b200 import capi
    from dealiiscale_b200.micro import MicroSolver, unique_id
    from dealiiscale_b200.parallel import MacroSharding
    from dealiiscale_b200.prm import BioData

    n_gpus = world
    micro_cells, macro_cells = 2 ** args.micro_ref, 2 ** args.macro_ref
    prm_path = os.path.join(ROOT, "input", args.prm)
    data = BioData(prm_path)
    xy_all = macro_points(macro_cells, n_gpus, strong=args.scaling == "strong")
    n_total = xy_all.shape[0]
    sharding = MacroSharding(n_total, rank, n_gpus)  # contiguous blocks of macro points, as msgpu_partition
    k0, k1 = sharding.k0, sharding.k1
    u_all, w_all = macro_values(xy_all)

    ms = MicroSolver(capi.BIO, data, micro_cells, device=local_rank)
    # run on torch's current stream so that torch.cuda.Event brackets exactly the launching stream
    ms.use_stream(torch.cuda.current_stream().cuda_stream)
    ms.precond = {"identity": capi.PRECOND_IDENTITY, "jacobi": capi.PRECOND_JACOBI, "ssor": capi.PRECOND_SSOR}[args.precond]
    ms.set_grid_locations(xy_all[k0:k1])
    if distributed:
        ids = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ms.attach_communicator(ids[0], rank, n_gpus, n_total)
    ms.setup()
    fused_exchange = distributed and not args.nccl_allgather
    if fused_exchange:
        # peer-memory exchange; if CUDA IPC is refused on any rank, every rank goes back to ncclAllGather together
        ok = 1
        try:
            if os.environ.get("MSGPU_BENCH_FAIL_P2P") == str(rank):  # test hook: this rank pretends CUDA IPC is refused
                raise capi.MsgpuError(capi.ERR_CUDA, "simulated")
            mine = ms.p2p_export()
        except capi.MsgpuError as err:
            mine, ok = b"\0" * 64, 0
            print(f"[bench] rank {rank}: p2p export failed ({err}); falling back to ncclAllGather", file=sys.stderr)
        handles = [None] * n_gpus
        dist.all_gather_object(handles, mine)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            try:
                ms.p2p_attach(handles)
            except capi.MsgpuError as err:
                ok = 0
                print(f"[bench] rank {rank}: p2p attach failed ({err}); falling back to ncclAllGather", file=sys.stderr)
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            ms.p2p_detach()
            fused_exchange = False
            args.nccl_allgather = True
    n = ms.n_dofs
    m = micro_cells + 1
    nnz = (3 * m - 2) ** 2

    # pinned host staging for the end-to-end path
    u_pin = torch.from_numpy(u_all.copy()).pin_memory()
    w_pin = torch.from_numpy(w_all.copy()).pin_memory()
    u_host, w_host = u_pin.numpy(), w_pin.numpy()

    def step_device():
        """inputs resident in HBM; results stay on the device"""
        ms.zero_solutions()
        ms.assemble_system()
        its = ms.solve()
        ms.coupling_on_device()
        return its

    def step_e2e():
        """through the public API with host buffers: H2D of the macro iterate, D2H of the integrals"""
        if distributed:
            ms.broadcast_macro(u_host, w_host, root=0)
        else:
            ms.set_macro_solution(u_host[k0:k1], w_host[k0:k1])
        ms.zero_solutions()
        ms.assemble_system()
        its = ms.solve()
        cu, cw = ms.coupling()
        return its, cu, cw

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    ms.set_macro_solution(u_all[k0:k1], w_all[k0:k1])
    for _ in range(max(3, args.warmup)):
        its = step_device()
    barrier()

    # ---- timed region 1: device-resident -----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms.enable_timing(True)
    ms.stats(reset=True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        its = step_device()
    ev1.record()
    barrier()
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    st = ms.stats(reset=True)
    ms.enable_timing(False)

    # ---- timed region 2: end to end through host buffers ----------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev2.record()
    for _ in range(args.steps):
        its_e, cu, cw = step_e2e()
    ev3.record()
    barrier()
    t_e2e = max(ev2.elapsed_time(ev3) * 1e-3, 0.0)
    t_e2e_wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- exchange check (N > 1): what this rank received for EVERY macro point against what the owner computed --------
    exchange_check = None
    if distributed:
        # an independent allgather of the owners' values through torch.distributed (NCCL) against what libmsgpu's
        # exchange (fused peer-memory kernel or ncclAllGather) delivered to this rank
        bad = sharding.check_exchange(cu, cw, local=(cu[k0:k1], cw[k0:k1]), device="cuda")
        exchange_check = "ok" if bad == 0 else f"{bad} entries differ"

    # device-side durations (CUDA events on the launching stream, recorded inside libmsgpu)
    dev_ms_step = (st["ms_tables"] + st["ms_moments"] + st["ms_gather"] + st["ms_cg"] + st["ms_coupling"]) / args.steps
    times = torch.tensor([t_dev, t_e2e, dev_ms_step, st["ms_cg"] / args.steps], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, dev_ms_step, cg_ms = [float(v) for v in times.tolist()]

    step_bytes, cg_bytes = algorithmic_bytes(n, nnz, its, n_total)
    if rank == 0:
        # FLOPs of one CG step: SpMV 2 nnz, two dot products 4 n, three vector updates 6 n (this rank's systems)
        cg_flops = float(its.sum()) * (2.0 * nnz + 10.0 * n)
        sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        fp64 = fp64_record(cg_flops, cg_ms, n_sms, sm_mhz)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = json.load(open(peaks_path))["hbm_gbs"]
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic, traffic_src, pipe_ncu = None, None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("cg_variant") == int(st["cg_variant"]):
                traffic, traffic_src = tj.get("cg_bytes_per_launch"), tj.get("source")
                pipe_ncu = tj.get("fp64_pipe_active_pct")  # same committed capture: not measured in this run
        achieved = cg_bytes / (cg_ms * 1e-3) / 1e9
        value = n_total * n * args.steps / t_dev
        e2e_value = n_total * n * args.steps / t_e2e
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n_gpus),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "wall_ms_per_step": 1e3 * t_e2e_wall / args.steps,
                    "h2d_bytes_per_step": int(16 * (n_total if distributed else (k1 - k0))),
                    "d2h_bytes_per_step": int(16 * n_total + 4 * (k1 - k0) + 16)},
            "gpu_launches": int(st["launches"]),
            "clocks": clocks,
            # The dominant kernel keeps every system ON CHIP for the whole solve (registers, tensor memory, shared
            # memory): it reads b and x once and writes x once, so HBM does not bound it - the FP64 pipe and the
            # latency of the two block-wide reductions per CG step do.  `frac` is therefore the FP64 fraction;
            # the SURVEY 8d byte model (which charges 8 nnz bytes per CG step) is kept under `hbm_model` for reference.
            "roofline": {
                "bound": "fp64/on-chip", "kernel": CG_KERNEL_NAMES.get(int(st["cg_variant"]), "batched CG"),
                "achieved": fp64["achieved"], "peak": fp64["peak"], "unit": "TFLOP/s", "frac": fp64["frac"],
                "traffic": traffic, "traffic_source": traffic_src,
                "fp64_pipe_active_pct_ncu": pipe_ncu,
                "peak_source": fp64["peak_source"], "flops_model": fp64["flops_model"],
                "algorithmic_flops_per_launch": cg_flops,
                "kernel_ms_per_launch": cg_ms,
                "hbm_model": {
                    "note": "SURVEY 8d byte model: 8 nnz bytes per CG step as if the matrix were streamed; the kernel does not "
                            "move these bytes (see traffic), so this 'fraction' can exceed 1 and is not a utilisation",
                    "algorithmic_bytes_per_launch": cg_bytes, "achieved_GBps": achieved, "peak_GBps": peak,
                    "peak_source": peak_src, "frac": achieved / peak, "frac_of_nominal_8TBps": achieved / 8000.0,
                    "step_algorithmic_bytes": step_bytes,
                    "step_achieved_GBps": step_bytes / (dev_ms_step * 1e-3) / 1e9,
                },
            },
            "phases_ms_per_step": {k: st[k] / args.steps for k in
                                   ("ms_tables", "ms_moments", "ms_gather", "ms_cg", "ms_coupling")},
            "cg_iterations": {"mean": float(its.mean()), "max": int(its.max()), "min": int(its.min())},
        }
        if exchange_check is not None:
            line["exchange_check"] = exchange_check
        if not args.no_cpu_baseline and n_gpus == 1:
            threads = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_baseline(prm_path, macro_cells, micro_cells, args.cpu_sample or 128 * threads,
                                                threads)
            # what solve_elliptic / solve_parabolic do in the reference: one thread over the grids
            one = cpu_baseline(prm_path, macro_cells, micro_cells, 49, 1)
            line["cpu_baseline"]["single_thread"] = {"value": one["value"], "unit": UNIT, "sample": one["sample"]}
    else:
        line = None
    ms.close()
    torch.cuda.synchronize()

    # ---- the other BASELINE configs (1 GPU): parabolic micro step, elliptic S2 with per-point matrices, ... -----------
    if rank == 0 and n_gpus == 1 and not args.no_other_configs:
        line["other_configs"] = other_configs(args, torch, local_rank, clocks)

    # ---- the real thing: the C++ driver solve_biomath on the literal lattice, all ranks, whole fixed-point cycles ----
    if not args.no_cycle:
        cyc = driver_cycle(args, rank, world, dist if distributed else None)
        if rank == 0 and cyc is not None:
            line["cycle"] = cyc["cycle"]
            line["strong"] = cyc["strong"]
        # the same driver with MSGPU_REUSE_ASSEMBLY=1 (opt-in): matrices and macro-independent loads are assembled in
        # the first cycle only - what the reference could do as well (only the right-hand sides depend on u, w);
        # reported beside the faithful loop, never instead of it
        cyc2 = driver_cycle(args, rank, world, dist if distributed else None, reuse_assembly=True)
        if rank == 0 and cyc2 is not None and "error" not in cyc2["cycle"]:
            c2 = cyc2["cycle"]
            line["cycle_with_assembly_reuse"] = {k: c2[k] for k in ("ms_macro", "ms_micro", "ms_stop_test", "ms_cycle",
                                                                   "mean_cg_steps", "later_cycles_ms_median")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
