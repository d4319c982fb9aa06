"""Graphs and parameter sets of the modularity-optimiser parity tests (shared by the golden
generator, the CPU emulation tests and the GPU tests).  Everything is seeded."""
import hashlib

import numpy as np
import scipy.sparse as sp

# the "karate club" network of /root/reference/tests/testthat/test_modularity_optimizer.R:9-20
KARATE_NODE1 = [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3,
                3, 4, 4, 5, 5, 5, 6, 8, 8, 8, 9, 13, 14, 14, 15, 15, 18, 18, 19, 20, 20, 22, 22, 23, 23, 23, 23, 23,
                24, 24, 24, 25, 26, 26, 27, 28, 28, 29, 29, 30, 30, 31, 31, 32]
KARATE_NODE2 = [1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 17, 19, 21, 31, 2, 3, 7, 13, 17, 19, 21, 30, 3, 7, 8, 9, 13,
                27, 28, 32, 7, 12, 13, 6, 10, 6, 10, 16, 16, 30, 32, 33, 33, 33, 32, 33, 32, 33, 32, 33, 33, 32, 33,
                32, 33, 25, 27, 29, 32, 33, 25, 27, 31, 31, 29, 33, 33, 31, 33, 32, 33, 32, 33, 32, 33, 33]

# known answers held by the reference's own tests (same file, :24-75): (modularityFunction, resolution,
# algorithm, nRandomStarts, nIterations, randomSeed) -> expected cluster vector
KARATE_KAT = [
    ((1, 1.0, 1, 1, 1, 564), [1, 1, 1, 1, 2, 2, 2, 1, 0, 1, 2, 1, 1, 1, 0, 0, 2, 1, 0, 1, 0, 1, 0, 0, 3, 3, 0, 0, 3, 0,
                              0, 3, 0, 0]),
    ((1, 1.0, 2, 1, 1, 2), [1, 1, 1, 1, 3, 3, 3, 1, 0, 0, 3, 1, 1, 1, 0, 0, 3, 1, 0, 1, 0, 1, 0, 2, 2, 2, 0, 2, 2, 0, 0,
                            2, 0, 0]),
]
# SLM known answers of the same file (:56-110)
KARATE_KAT_SLM = [
    ((1, 1.0, 3, 1, 1, 56464), [1, 1, 1, 1, 3, 3, 3, 1, 0, 0, 3, 1, 1, 1, 0, 0, 3, 1, 0, 1, 0, 1, 0, 2, 2, 2, 0, 2, 2,
                                0, 0, 2, 0, 0]),
    ((1, 0.05, 3, 1, 10, 10), [0] * 34),
    ((2, 0.05, 3, 1, 10, 10), [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1,
                               1, 1, 1, 1]),
]


def karate():
    """lower-triangular dgCMatrix, as the reference test builds it (sparseMatrix(i = node2 + 1, j = node1 + 1))"""
    m = sp.csc_matrix((np.ones(len(KARATE_NODE1)), (KARATE_NODE2, KARATE_NODE1)), shape=(34, 34))
    m.sort_indices()
    return m


def karate_weighted():
    """the EdgeWeights case of the reference test (:112-121)"""
    m = karate().tolil()
    for (r, c, w) in ((5, 4, 3.0), (5, 1, 5.0), (4, 1, 8.0), (20, 5, 8.0), (20, 4, 5.0), (20, 1, 5.0)):
        m[r - 1, c - 1] = w
    m = m.tocsc()
    m.sort_indices()
    return m


def random_graph(n, avg_deg, seed, weighted=True, symmetric=True, isolated=0):
    """seeded sparse graph with a planted partition; dgCMatrix slots.  `isolated` trailing nodes have no edge."""
    rng = np.random.default_rng(seed)
    m = n - isolated
    groups = rng.integers(0, max(2, m // 40), m)
    e = int(m * avg_deg / 2)
    a = rng.integers(0, m, 3 * e)
    b = rng.integers(0, m, 3 * e)
    keep = (a != b) & ((groups[a] == groups[b]) | (rng.random(3 * e) < 0.15))
    a, b = a[keep][:e], b[keep][:e]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = np.unique(lo.astype(np.int64) * n + hi)
    lo, hi = (key // n).astype(np.int64), (key % n).astype(np.int64)
    w = rng.integers(1, 16, lo.size) / rng.integers(1, 24, lo.size) if weighted else np.ones(lo.size)
    if symmetric:
        rows, cols, vals = np.concatenate([hi, lo]), np.concatenate([lo, hi]), np.concatenate([w, w])
    else:
        rows, cols, vals = hi, lo, w      # lower triangle only, like the reference test's matrix
    M = sp.csc_matrix((vals, (rows, cols)), shape=(n, n))
    M.sort_indices()
    return M


def slots(M):
    return M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)


def digest(labels) -> str:
    return hashlib.sha256(np.ascontiguousarray(labels, dtype=np.int32).tobytes()).hexdigest()


# (name, builder, [(modularityFunction, resolution, algorithm, nRandomStarts, nIterations, seed), ...])
def small_cases():
    return [
        ("karate", karate, [(1, 1.0, 1, 1, 1, 564), (1, 1.0, 2, 1, 1, 2), (1, 0.05, 1, 1, 10, 10), (2, 0.05, 2, 1, 10, 10),
                            (1, 1.0, 1, 10, 10, 0), (1, 0.5, 2, 5, 3, 7), (2, 1.0, 1, 3, 4, 99), (1, 3.0, 1, 2, 2, 5),
                            (1, 1.0, 3, 1, 1, 56464), (1, 0.05, 3, 1, 10, 10), (2, 0.05, 3, 1, 10, 10), (1, 1.3, 3, 4, 3, 77)]),
        ("karate_weighted", karate_weighted, [(1, 1.0, 1, 1, 10, 40), (1, 1.0, 2, 1, 10, 40), (2, 0.3, 1, 4, 5, 1),
                                              (1, 1.0, 3, 1, 10, 40)]),
        ("rand300_w_sym", lambda: random_graph(300, 8, 1), [(1, 0.8, 1, 3, 10, 0), (1, 1.5, 2, 2, 5, 11),
                                                            (2, 0.02, 1, 2, 3, 4), (1, 25.0, 1, 1, 2, 3),
                                                            (1, 0.8, 3, 2, 4, 9), (2, 0.05, 3, 1, 2, 6)]),
        ("rand1000_unw_lower_iso", lambda: random_graph(1000, 6, 2, weighted=False, symmetric=False, isolated=7),
         [(1, 1.0, 1, 2, 10, 123456789), (1, 0.3, 2, 1, 10, 8), (2, 0.004, 2, 2, 2, 2), (1, 1.0, 3, 2, 3, 31)]),
        ("rand5000_w_sym", lambda: random_graph(5000, 14, 3), [(1, 0.8, 1, 2, 10, 0), (1, 2.0, 2, 1, 4, 17),
                                                               (1, 0.8, 3, 1, 3, 5)]),
    ]
