"""CPU tier: the parts of bench.py's contract that need no GPU.

* `--impl reference` (the CPU arm: the oracle on the host's threads) prints ONE JSON line with the keys the driver
  reads, on the same workload string as the GPU arm;
* the committed final line of the round (profiles/r2_final_bench_1gpu.json) carries every key of the contract, a
  roofline fraction <= 1 and the copies declared in `e2e`."""
import json
import os
import subprocess
import sys

from helpers import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config"]


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED + ["impl", "cpu_baseline", "e2e"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["metric"] == "micro_dof_solves_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "biocase.prm" in d["config"]["workload"]


def test_committed_gpu_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r2_final_bench_1gpu.json")
    d = json.load(open(path))
    for k in REQUIRED + ["e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "other_configs", "cycle", "strong"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    r = d["roofline"]
    assert 0 < r["frac"] <= 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and r["traffic"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] <= d["value"] * 1.02        # the end-to-end figure does not just repeat the device one
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert len(d["other_configs"]) == 4
