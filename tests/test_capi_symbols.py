"""The C-ABI library loads on a machine without a GPU and exports every symbol include/msgpu.h
declares; nothing is computed here."""
import ctypes
import os
import re

from helpers import ROOT
from dealiiscale_b200 import capi


def _declared():
    text = open(os.path.join(ROOT, "include", "msgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msgpu_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTED) == declared


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to create a context."""
    import torch

    if torch.cuda.is_available():
        return
    lib = capi.load()
    h = ctypes.c_void_p()
    desc = capi.Desc(capi.ELLIPTIC, 0, 4, 12, -1)
    assert lib.msgpu_create(ctypes.byref(h), ctypes.byref(desc)) == capi.ERR_CUDA
    assert not h.value


def test_partition_matches_c():
    lib = capi.load()
    for n_total, ranks in ((4225, 8), (10, 3), (7, 8), (33800, 8), (1, 1)):
        for r in range(ranks):
            k0, k1 = ctypes.c_int(), ctypes.c_int()
            assert lib.msgpu_partition(n_total, ranks, r, ctypes.byref(k0), ctypes.byref(k1)) == 0
            assert (k0.value, k1.value) == capi.partition(n_total, ranks, r)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "dealiiscale_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f == "build.py":  # building the checker is not using it
                continue
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                for needle in ("liboracle", "oracle/_build", "oracle/_ref", 'include "../../oracle', "import oracle",
                               "from oracle", "load_oracle", "helpers import"):
                    assert needle not in text, (f, needle)


def test_cmake_build_lists_the_same_sources_as_the_in_tree_recipe():
    """CMakeLists.txt (for a CMake host project) and dealiiscale_b200/build.py (tests, bench) build the same library."""
    from dealiiscale_b200 import build

    text = open(os.path.join(ROOT, "CMakeLists.txt")).read()
    for src in build.CUDA_SOURCES + build.HOST_SOURCES:
        assert "${CSRC}/" + src in text, src
    for exe in ("solve_elliptic", "solve_biomath", "solve_parabolic"):
        assert exe in text
    assert "CMAKE_CUDA_ARCHITECTURES 100a" in text
