"""CPU tier: the bounds-asserting build (csrc/checked.cuh) is a separate library and leaves no trace in the product.

`libmsgpu_checked.so` carries the assert messages of MSGPU_CHECK; `libmsgpu.so` must not (the checks expand to nothing
there), and the register budgets that decide how many CTAs of the resident CG kernels share an SM must be those the
launch code counts on."""
import os
import re
import shutil
import subprocess

import pytest

from helpers import ROOT

BUILD = os.path.join(ROOT, "dealiiscale_b200", "_build")
MARK = b"[msgpu checked build]"


def _libs():
    import sys
    sys.path.insert(0, ROOT)
    from dealiiscale_b200 import build as b
    return b.build_msgpu(), b.build_checked()


def test_checks_exist_only_in_the_checked_library():
    product, checked = _libs()
    assert MARK not in open(product, "rb").read()
    assert MARK in open(checked, "rb").read()
    # nothing in the product refers to the checked library
    for root, _, files in os.walk(os.path.join(ROOT, "dealiiscale_b200")):
        if "_build" in root or "__pycache__" in root:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")) and f not in ("build.py", "checked.cuh"):
                assert "libmsgpu_checked" not in open(os.path.join(root, f), errors="ignore").read(), f


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_register_budgets_of_the_resident_cg_kernels():
    product, _ = _libs()
    out = subprocess.run(["cuobjdump", "-res-usage", product], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    stencil_tm = [r for n, r in regs.items() if "k_cg_stencilILb1" in n]
    assert stencil_tm and max(stencil_tm) <= 128        # two CTAs of 256 threads per SM
    tmem17 = [r for n, r in regs.items() if "k_cg_strip_tmemILi256ELi17" in n]
    assert tmem17 and max(tmem17) <= 255                 # one CTA of 256 threads per SM
    tmem9 = [r for n, r in regs.items() if "k_cg_strip_tmemILi256ELi9" in n]
    assert tmem9 and max(tmem9) <= 128                   # two per SM (mid-size systems)
