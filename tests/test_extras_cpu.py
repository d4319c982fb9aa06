"""CPU suite of the SNN side consumers: the oracle's restatements of WriteEdgeFile,
fast_dist and SNN_SmallestNonzero_Dist against independent Python restatements of the reference
lines, the host emulation of the device row bodies (seurat_b200/csrc/extras_core.inl) against the
oracle, and the edge-file writer of the library (host I/O, needs no GPU) against the oracle's."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import scipy.sparse as sp

import oracle
import seurat_b200 as sb
from seurat_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))


def emul():
    src = os.path.join(HERE, "emul", "extras_emul.cpp")
    core = os.path.join(os.path.dirname(HERE), "seurat_b200", "csrc", "extras_core.inl")
    lib = os.path.join(HERE, "emul", "libextras_emul.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-w", src,
                               "-o", lib])
    return C.CDLL(lib)


def make_snn(n=400, d=7, k=12, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    X[5] = X[4]                                   # a duplicated cell: zero distances
    r = oracle.find_neighbors(X, k=k, prune=1 / 15)
    return X, r


def py_smallest(p, i, x, mat, n, nearest):
    """src/snn.cpp:82-126 line by line"""
    out = []
    for col in range(len(p) - 1):
        vals = x[p[col]:p[col + 1]]
        rows = i[p[col]:p[col + 1]]
        order = sorted(range(len(vals)), key=lambda q: vals[q])       # stable
        n_i = min(n, len(order))
        dists = []
        for j in range(len(order)):
            if len(dists) < n_i or vals[order[j]] == vals[order[n_i - 1]]:
                res = math.sqrt(sum((a - b) ** 2 for a, b in zip(mat[rows[order[j]]], mat[col])))
                if nearest[col] > 0:
                    res = max(res - nearest[col], 0.0)
                dists.append(res)
            else:
                break
        if len(dists) > n_i:
            dists.sort(reverse=True)
            out.append(sum(dists[:n_i]) / n_i)
        else:
            out.append(sum(dists) / len(dists) if dists else float("nan"))
    return np.array(out)


def test_edge_file_writer_matches_the_oracle_and_the_format(tmp_path):
    X, r = make_snn()
    p, i, x = r["snn"]
    f1, f2 = str(tmp_path / "lib.txt"), str(tmp_path / "oracle.txt")
    n1 = _lib.write_edge_file(p, i, x, f1)
    n2 = oracle.write_edge_file(p, i, x, f2)
    S = sp.csc_matrix((x, i, p), shape=(len(p) - 1,) * 2)
    assert n1 == n2 == sp.tril(S, -1).nnz
    assert open(f1, "rb").read() == open(f2, "rb").read()
    first = open(f1).readline().rstrip("\n").split("\t")
    col, row, val = int(first[0]), int(first[1]), first[2]
    assert col < row and val == "%.15g" % S[row, col]
    # the mirror of the R-level helpers
    g = sb.Graph(p, i, x, S.shape)
    f3 = str(tmp_path / "mirror.txt")
    sb.WriteEdgeFile(g, f3)
    assert open(f3, "rb").read() == open(f1, "rb").read()
    # arbitrary doubles and an empty matrix
    M = sp.random(50, 50, 0.2, random_state=3, format="csc")
    M.sort_indices()
    f4, f5 = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    _lib.write_edge_file(M.indptr, M.indices, M.data, f4)
    oracle.write_edge_file(M.indptr, M.indices, M.data, f5)
    assert open(f4, "rb").read() == open(f5, "rb").read()
    assert _lib.write_edge_file(np.zeros(4, np.int32), np.zeros(0, np.int32), np.zeros(0), f4) == 0


def test_fast_dist_oracle_and_emulation():
    rng = np.random.default_rng(1)
    x, y = rng.normal(size=(60, 9)), rng.normal(size=(80, 9))
    lists = [rng.integers(1, 81, rng.integers(0, 7)) for _ in range(60)]
    got = oracle.fast_dist(x, y, lists)
    for r in (0, 17, 59):
        for q, j in enumerate(lists[r]):
            acc = 0.0
            for c in range(9):
                acc += (x[r, c] - y[j - 1, c]) ** 2
            assert got[r][q] == math.sqrt(acc)
    ptr = np.zeros(61, np.int64)
    ptr[1:] = np.cumsum([len(v) for v in lists])
    nbr = np.concatenate(lists).astype(np.int32)
    out = np.empty(int(ptr[-1]))
    emul().emul_fast_dist(C.c_void_p(x.ctypes.data), C.c_int64(60), C.c_void_p(y.ctypes.data), 9,
                          C.c_void_p(ptr.ctypes.data), C.c_void_p(nbr.ctypes.data), C.c_void_p(out.ctypes.data))
    assert (out == np.concatenate(got)).all()


def test_smallest_nonzero_dist_oracle_and_emulation():
    X, r = make_snn()
    p, i, x = r["snn"]
    nearest = r["nn_dist"][:, 1].copy()
    nearest[::7] = 0.0                               # nearest_dist == 0: no subtraction
    for n_small in (1, 5, 20, 10_000):
        want = oracle.snn_smallest_nonzero_dist(p, i, x, X, n_small, nearest)
        ref = py_smallest(p, i, x, X, n_small, nearest)
        assert np.allclose(want, ref, rtol=1e-13, atol=0, equal_nan=True)
        out = np.empty(len(X))
        p64 = p.astype(np.int64)
        emul().emul_smallest_nonzero(C.c_void_p(p64.ctypes.data), C.c_void_p(i.ctypes.data), C.c_void_p(x.ctypes.data),
                                     C.c_int64(len(X)), C.c_void_p(X.ctypes.data), X.shape[1], n_small,
                                     C.c_void_p(nearest.ctypes.data), C.c_void_p(out.ctypes.data))
        assert (out == want).all()
    # a column without entries gives NaN, as the reference's 0 / 0 does
    pe = np.array([0, 0, 1, 2], np.int32)
    w = oracle.snn_smallest_nonzero_dist(pe, np.array([2, 1], np.int32), np.array([0.5, 0.5]), X[:3], 2, np.zeros(3))
    assert np.isnan(w[0]) and np.isfinite(w[1:]).all()
