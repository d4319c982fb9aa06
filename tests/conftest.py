import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fixtures():
    with open(os.path.join(ROOT, "tests", "golden", "fixtures.json")) as f:
        return json.load(f)


def unhex(v):
    """fixtures store doubles as C99 hex strings so that they round-trip exactly
    (either a plain list, or a table of distinct values plus indices)"""
    if isinstance(v, dict):
        table = np.array([float.fromhex(s) for s in v["vals"]], dtype=np.float64)
        return table[np.array(v["idx"], dtype=np.int64)] if len(v["idx"]) else np.empty(0)
    return np.array([float.fromhex(s) if isinstance(s, str) else float(s) for s in v], dtype=np.float64)


def unhex2(rows):
    return np.array([[float.fromhex(s) for s in r] for r in rows], dtype=np.float64)


@pytest.fixture(scope="session")
def pbmc(fixtures):
    f = fixtures["pbmc80x10"]
    return {
        "cells": f["cells"],
        "X": unhex2(f["embedding"]),
        "k": f["k"],
        "prune": float.fromhex(f["prune"]),
        "nn_idx": np.array(f["nn_idx"], dtype=np.int32),
        "nn_dist": unhex2(f["nn_dist"]),
        "nn": (np.array(f["nn"]["p"]), np.array(f["nn"]["i"]), unhex(f["nn"]["x"])),
        "snn": (np.array(f["snn"]["p"]), np.array(f["snn"]["i"]), unhex(f["snn"]["x"])),
    }


def assert_csr_equal(got, want, what=""):
    """bit-exact comparison of (p, i, x) triples"""
    gp, gi, gx = (np.asarray(a) for a in got)
    wp, wi, wx = (np.asarray(a) for a in want)
    assert gp.shape == wp.shape and (gp.astype(np.int64) == wp.astype(np.int64)).all(), f"{what}: p differs"
    assert gi.shape == wi.shape and (gi.astype(np.int64) == wi.astype(np.int64)).all(), f"{what}: i differs"
    assert gx.shape == wx.shape, f"{what}: x length differs"
    # bitwise (inf == inf, and there must be no NaN)
    assert (gx.view(np.int64) == wx.view(np.int64)).all(), f"{what}: x differs"


def assert_knn_matches(idx, dist, ref_idx, ref_dist, what=""):
    """north_star's rule: indices identical except where the oracle's distances
    tie within 1e-6 relative (then the sets must agree); distances within 1e-5 relative."""
    idx, ref_idx = np.asarray(idx), np.asarray(ref_idx)
    dist, ref_dist = np.asarray(dist), np.asarray(ref_dist)
    assert idx.shape == ref_idx.shape and dist.shape == ref_dist.shape, what
    rel = np.abs(dist - ref_dist) <= 1e-5 * np.maximum(np.abs(ref_dist), 1e-300)
    assert (rel | (dist == ref_dist)).all(), f"{what}: distances differ by more than 1e-5 relative"
    bad_rows = np.nonzero((idx != ref_idx).any(axis=1))[0]
    for r in bad_rows:
        cols = np.nonzero(idx[r] != ref_idx[r])[0]
        for c in cols:
            # the mismatching position must sit inside a run of (near-)tied oracle distances
            dref = ref_dist[r, c]
            tied = np.abs(ref_dist[r] - dref) <= 1e-6 * max(abs(dref), 1e-300)
            assert tied.sum() > 1, f"{what}: row {r} col {c}: index {idx[r, c]} != {ref_idx[r, c]} without a tie"
            assert set(idx[r][tied]) == set(ref_idx[r][tied]) or c == idx.shape[1] - 1 or tied[-1], \
                f"{what}: row {r}: tied run holds different neighbours"
