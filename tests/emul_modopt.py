"""Loader of the serial emulation of the device modularity optimiser (tests/emul/modopt_emul.cpp):
the step functions of seurat_b200/csrc/modopt_core.inl, compiled for the host.  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "modopt_emul.cpp")
CORE = os.path.join(os.path.dirname(HERE), "seurat_b200", "csrc", "modopt_core.inl")
LIB = os.path.join(HERE, "emul", "libmodopt_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-w",
                                   SRC, "-o", LIB])
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.emul_modularity_cluster.argtypes = [C.c_int, vp, vp, vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                              C.c_uint64, vp, vp, C.c_int64, C.c_int64, vp]
        _lib = L
    return _lib


def cluster(p, i, x, mf, res, alg, ns, ni, seed, win0=0, winmax=0):
    """-> (labels, n_clusters, modularity, stats dict)"""
    p = np.ascontiguousarray(p, dtype=np.int32)
    i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = p.size - 1
    out = np.zeros(n, np.int32)
    q = C.c_double()
    st = np.zeros(16, np.int64)
    rc = lib().emul_modularity_cluster(n, p.ctypes.data, i.ctypes.data, x.ctypes.data, mf, float(res), alg, ns, ni,
                                       seed, out.ctypes.data, C.addressof(q), win0, winmax, st.ctypes.data)
    names = ("windows", "rounds", "sequential_steps", "window_retries", "local_moving_calls", "levels", "max_window")
    return out, rc, q.value, dict(zip(names, st.tolist()))


def cluster_edges(node1, node2, w, n, mf, res, alg, ns, ni, seed):
    """-> (labels, n_clusters, modularity) from an edge list (build_network_edges of modopt_core.inl)"""
    node1 = np.ascontiguousarray(node1, dtype=np.int32)
    node2 = np.ascontiguousarray(node2, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.zeros(n, np.int32)
    q = C.c_double()
    L = lib()
    vp = C.c_void_p
    L.emul_modularity_cluster_edges.argtypes = [C.c_int, vp, vp, vp, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int,
                                                C.c_int, C.c_uint64, vp, vp]
    rc = L.emul_modularity_cluster_edges(n, node1.ctypes.data, node2.ctypes.data, w.ctypes.data, node1.size, mf,
                                         float(res), alg, ns, ni, seed, out.ctypes.data, C.addressof(q))
    return out, rc, q.value


def seq_sum(x, init=0.0):
    """the order-fixed running sum of modopt_core.inl (binade-wise integer prefix sums)"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    L = lib()
    L.emul_seq_sum.argtypes = [C.c_void_p, C.c_int64, C.c_double]
    L.emul_seq_sum.restype = C.c_double
    return float(L.emul_seq_sum(x.ctypes.data, x.size, float(init)))
