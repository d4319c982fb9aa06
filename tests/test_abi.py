"""CPU suite: the C-ABI library builds, loads and exports every symbol that
include/b200nn.h declares; without a GPU every compute entry point fails loudly
(there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import seurat_b200
from seurat_b200 import _lib, build as b200_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200nn.h")).read()
    return sorted(set(re.findall(r"B200NN_API[^;]*?\b(b200nn_\w+)\s*\(", src)))


def test_library_builds_and_exports_all_header_symbols():
    path = b200_build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"{s} declared in b200nn.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms, "the ctypes binding must cover exactly the header's entry points"
    assert L.b200nn_version() == 2


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "b200nn.h")).read()
    assert "torch" not in src.lower() and "at::" not in src and "#include <cuda" not in src


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    with pytest.raises(seurat_b200.B200Error) as ei:
        seurat_b200.Context(0)
    assert ei.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)
    X = seurat_b200.Embedding(np.random.default_rng(0).normal(size=(30, 4)),
                              [f"c{i}" for i in range(30)])
    with pytest.raises(seurat_b200.B200Error):
        seurat_b200.FindNeighbors(X, k_param=5, verbose=False)
    with pytest.raises(seurat_b200.B200Error):
        seurat_b200.ComputeSNN(np.ones((4, 2), dtype=np.int32), 0.0)


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under seurat_b200/ or the shim may reference it"""
    for base in ("seurat_b200", "shim", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".R")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "liboracle" not in txt and "oracle_" not in txt, \
                        f"{os.path.join(dirpath, f)} references the oracle"


def test_r_shim_compiles_against_the_documented_r_api():
    """R is not in the image, so the .Call shim cannot be built for real; it is at least compiled
    (gcc -fsyntax-only -Wall -Werror) against declarations of the R API entry points it uses
    (tests/stubs/r_api, written from "Writing R Extensions") and the real include/b200nn.h: argument
    counts and types of every R API and every C-ABI call are checked, and the registration table must
    list every .Call entry point with its arity (src/RcppExports.cpp:407-445 pattern)."""
    import subprocess
    shim = os.path.join(ROOT, "shim", "b200nn_r.c")
    r = subprocess.run(["/usr/bin/gcc", "-std=gnu11", "-fsyntax-only", "-Wall", "-Werror",
                        "-I" + os.path.join(ROOT, "tests", "stubs", "r_api"), "-I" + os.path.join(ROOT, "include"),
                        shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(shim).read()
    defined = {m.group(1): m.group(2).count("SEXP") for m in re.finditer(r"^SEXP (b200_\w+)\(([^)]*)\)\s*\{", src, re.M)}
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(b200_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', src)}
    assert defined == table, "every .Call entry point is registered with its number of arguments"
    rsrc = open(os.path.join(ROOT, "shim", "b200nn.R")).read()
    for name in re.findall(r'\.Call\("(b200_\w+)"', rsrc):
        assert name in table, f"{name} is called from b200nn.R but not registered"


def test_every_entry_point_is_documented():
    """INTEGRATION.md lists the whole ABI: a new entry point must be explained to the maintainer who binds it"""
    docs = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    listed = set(re.findall(r"`(b200nn_\w+)`", docs))
    # "`b200nn_x_count` / `_fill`" style continuations
    for stem, tails in re.findall(r"`(b200nn_\w+?)_(?:count|create)`((?:\s*/\s*`_\w+`)+)", docs):
        for t in re.findall(r"`_(\w+)`", tails):
            listed.add(f"{stem}_{t}")
    missing = [s for s in header_symbols() if s not in listed]
    assert not missing, f"not mentioned in INTEGRATION.md: {missing}"
