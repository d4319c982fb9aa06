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


def test_integer_restatement_of_the_long_double_sum_is_exact():
    """X87Sum (extras_core.inl) == the compiler's own long double accumulation (x87 extended on x86-64, what R's
    sum() uses), bit for bit after the final rounding to double: random magnitudes over 600 binades, long
    sums, ties, subnormals, zeros, huge terms; then L2Norm rows against the oracle's L2Norm"""
    L = emul()
    assert L.native_long_double_mantissa_bits() == 64, "the pin needs an x87 long double (x86-64)"
    for f in (L.emul_x87_sum, L.native_long_double_sum):
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_int64]
    rng = np.random.default_rng(17)

    def both(t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        return L.emul_x87_sum(t.ctypes.data, t.size), L.native_long_double_sum(t.ctypes.data, t.size)

    for trial in range(4000):
        n = int(rng.integers(1, 120))
        kind = trial % 8
        if kind == 0:
            t = rng.random(n) ** 2
        elif kind == 1:
            t = (rng.normal(size=n) * 10.0 ** rng.integers(-150, 150, n)) ** 2
        elif kind == 2:                       # exact ties at the 64-bit rounding position
            t = np.ldexp(rng.integers(1, 2 ** 53, n).astype(np.float64), rng.integers(-70, 12, n))
        elif kind == 3:                       # subnormal and tiny terms
            t = np.ldexp(rng.integers(0, 2 ** 20, n).astype(np.float64), rng.integers(-1074, -1000, n))
        elif kind == 4:
            t = np.where(rng.random(n) < 0.3, 0.0, rng.random(n) * 1e-3)
        elif kind == 5:                       # one dominating term, many that only leave a sticky bit
            t = rng.random(n) * 1e-25
            t[rng.integers(0, n)] = 1e12
        elif kind == 6:
            t = np.full(n, np.ldexp(1.0, int(rng.integers(-40, 40)))) * (1 + 2.0 ** -52)
        else:
            t = np.ldexp(rng.random(n) + 0.5, rng.integers(900, 1023, n))       # near overflow
        a, b = both(t)
        assert a == b or (np.isinf(a) and np.isinf(b)), (trial, kind, a, b)
    big = rng.random(50_000) * 10.0 ** rng.integers(-8, 8, 50_000)
    a, b = both(big)
    assert a == b
    assert np.isinf(both(np.array([1.0, np.inf]))[0])
    # rows: the emulated device body == the oracle's L2Norm (long double in C), zero rows -> zeros
    L.emul_l2norm.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    for d in (1, 3, 10, 50, 99):
        M = rng.normal(size=(2000, d)) * 10.0 ** rng.integers(-6, 6, size=(2000, 1))
        M[7] = 0.0
        out = np.empty_like(M)
        L.emul_l2norm(M.ctypes.data, M.shape[0], d, out.ctypes.data)
        want = oracle.l2norm(M)
        assert (out == want).all() and (out[7] == 0).all()
