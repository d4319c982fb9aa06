"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  The product package
``seurat_b200`` never imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` do.

Parity status: kNN / SNN unpinned against the real reference (no R / RANN / Eigen in
the container); pinned against the SURVEY.md 8c known-answer vectors in
``tests/golden`` and cross-checked against scipy / sklearn in ``tests/``.

The modularity optimiser is different: ``modularity_cluster_ref`` runs the REAL reference
(``/root/reference/src/ModularityOptimizer.cpp`` compiled where it lies into
``oracle/_ref/libmodopt_ref.so`` by ``make -C oracle ref``) and is pinned by the known answers
the reference's own tests hold (``tests/testthat/test_modularity_optimizer.R``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i64, i32, dbl, vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
        L.oracle_knn_brute.argtypes = [vp, i64, vp, i64, i32, i32, vp, vp, i32]
        L.oracle_knn_brute.restype = i32
        L.oracle_knn_kdtree.argtypes = [vp, i64, vp, i64, i32, i32, vp, vp, i32, vp, vp]
        L.oracle_knn_kdtree.restype = i32
        L.oracle_nn_graph.argtypes = [vp, i64, i32, i64, vp, vp, vp, vp]
        L.oracle_nn_graph.restype = i32
        L.oracle_compute_snn.argtypes = [vp, i64, i32, dbl, i32, vp, vp]
        L.oracle_compute_snn.restype = i32
        L.oracle_compute_snn_rows.argtypes = [vp, i64, i32, dbl, i64, i64, i32, vp, vp, vp]
        L.oracle_compute_snn_rows.restype = i32
        L.oracle_snn_fetch.argtypes = [vp, vp, vp, vp]
        L.oracle_snn_fetch.restype = None
        L.oracle_l2norm.argtypes = [vp, i64, i32]
        L.oracle_l2norm.restype = None
        L.oracle_num_threads.restype = i32
        L.oracle_write_edge_file.argtypes = [vp, vp, vp, i64, C.c_char_p]
        L.oracle_write_edge_file.restype = i64
        L.oracle_fast_dist.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp]
        L.oracle_fast_dist.restype = i32
        L.oracle_snn_smallest_nonzero_dist.argtypes = [vp, vp, vp, i64, vp, i32, i32, vp, vp]
        L.oracle_snn_smallest_nonzero_dist.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _mat(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def knn(data, query=None, k: int = 20, method: str = "brute", nthreads: int = 0,
        timings: dict | None = None):
    """Exact kNN as ``RANN::nn2(data, query, k)`` returns it.

    Returns ``(nn_idx int32 [nq,k] 1-based, nn_dists float64 [nq,k])``.
    ``method`` is "brute" (ties by lower index; the parity reference) or
    "kdtree" (ANN-style, the timed CPU baseline).
    """
    ref = _mat(data)
    qry = ref if query is None else _mat(query)
    nr, d = ref.shape
    nq = qry.shape[0]
    if qry.shape[1] != d:
        raise ValueError("query and data must have the same number of columns")
    idx = np.empty((nq, k), dtype=np.int32)
    dist = np.empty((nq, k), dtype=np.float64)
    if method == "brute":
        rc = lib().oracle_knn_brute(_p(ref), nr, _p(qry), nq, d, k, _p(idx), _p(dist), nthreads)
    elif method == "kdtree":
        tb, ts = C.c_double(0), C.c_double(0)
        rc = lib().oracle_knn_kdtree(_p(ref), nr, _p(qry), nq, d, k, _p(idx), _p(dist), nthreads,
                                     C.byref(tb), C.byref(ts))
        if timings is not None:
            timings["build_s"], timings["search_s"] = tb.value, ts.value
    else:
        raise ValueError(method)
    if rc == 1:
        raise ValueError("Cannot find more nearest neighbours than there are points")
    if rc:
        raise RuntimeError(f"oracle kNN failed rc={rc}")
    return idx, dist


def nn_graph(nn_idx, n: int | None = None):
    """nn Graph of R/clustering.R:664-670 as dgCMatrix arrays (p, i, x)."""
    nn = np.ascontiguousarray(np.asarray(nn_idx, dtype=np.int32))
    rows, k = nn.shape
    n = rows if n is None else n
    nnz = C.c_int64(0)
    rc = lib().oracle_nn_graph(_p(nn), rows, k, n, C.byref(nnz), None, None, None)
    if rc:
        raise ValueError("neighbour index out of range")
    p = np.empty(n + 1, dtype=np.int32)
    i = np.empty(nnz.value, dtype=np.int32)
    x = np.empty(nnz.value, dtype=np.float64)
    lib().oracle_nn_graph(_p(nn), rows, k, n, C.byref(nnz), _p(p), _p(i), _p(x))
    return p, i, x


def compute_snn(nn_ranked, prune: float, nthreads: int = 0):
    """``ComputeSNN(nn_ranked, prune)`` (src/snn.cpp:14-38) -> (p int64, i int32, x float64)."""
    nn = np.ascontiguousarray(np.asarray(nn_ranked, dtype=np.int32))
    n, k = nn.shape
    h = C.c_void_p()
    nnz = C.c_int64(0)
    rc = lib().oracle_compute_snn(_p(nn), n, k, float(prune), nthreads, C.byref(h), C.byref(nnz))
    if rc:
        raise ValueError("neighbour index out of range")
    p = np.empty(n + 1, dtype=np.int64)
    i = np.empty(nnz.value, dtype=np.int32)
    x = np.empty(nnz.value, dtype=np.float64)
    lib().oracle_snn_fetch(h, _p(p), _p(i), _p(x))
    return p, i, x


def compute_snn_rows(nn_ranked, prune: float, row_begin: int, row_end: int, nthreads: int = 0,
                     timings: dict | None = None):
    """SNN rows [row_begin, row_end) only (local offsets); used for bounded CPU baselines."""
    nn = np.ascontiguousarray(np.asarray(nn_ranked, dtype=np.int32))
    n, k = nn.shape
    h = C.c_void_p()
    nnz = C.c_int64(0)
    sec = (C.c_double * 2)()
    rc = lib().oracle_compute_snn_rows(_p(nn), n, k, float(prune), row_begin, row_end, nthreads,
                                       C.byref(h), C.byref(nnz), sec)
    if rc:
        raise ValueError("neighbour index out of range")
    rows = row_end - row_begin
    p = np.empty(rows + 1, dtype=np.int64)
    i = np.empty(nnz.value, dtype=np.int32)
    x = np.empty(nnz.value, dtype=np.float64)
    lib().oracle_snn_fetch(h, _p(p), _p(i), _p(x))
    if timings is not None:
        timings["transpose_s"], timings["rows_s"] = sec[0], sec[1]
    return p, i, x


def l2norm(mat) -> np.ndarray:
    m = _mat(mat).copy()
    lib().oracle_l2norm(_p(m), m.shape[0], m.shape[1])
    return m


def find_neighbors(data, k: int = 20, prune: float = 1.0 / 15.0, nthreads: int = 0,
                   method: str = "brute"):
    """kNN + nn Graph + SNN on one matrix: what FindNeighbors.default computes."""
    idx, dist = knn(data, None, k, method=method, nthreads=nthreads)
    return {"nn_idx": idx, "nn_dist": dist, "nn": nn_graph(idx), "snn": compute_snn(idx, prune, nthreads)}


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ----------------------------------------------------------------------------------------------
# modularity optimiser: the real reference (oracle/_ref), see modopt_ref_driver.cpp
# ----------------------------------------------------------------------------------------------
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libmodopt_ref.so")
_ref_lib = None


def build_ref(force: bool = False) -> str | None:
    """Compile oracle/_ref/libmodopt_ref.so from the reference sources where they lie.  Returns the
    path, or None when /root/reference is absent (the GPU box: the prebuilt file travels there)."""
    src = "/root/reference/src/ModularityOptimizer.cpp"
    if os.path.exists(src) and (force or not os.path.exists(_REF_LIB_PATH)
                                or os.path.getmtime(_REF_LIB_PATH) < os.path.getmtime(
                                    os.path.join(_HERE, "modopt_ref_driver.cpp"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "ref"], stdout=subprocess.DEVNULL)
    return _REF_LIB_PATH if os.path.exists(_REF_LIB_PATH) else None


def have_ref() -> bool:
    return build_ref() is not None


def modularity_cluster_ref(p, i, x, modularity_function=1, resolution=0.8, algorithm=1, n_random_starts=10,
                           n_iterations=10, random_seed=0):
    """RunModularityClusteringCpp of the reference itself on dgCMatrix slots (p, i, x).
    Returns (cluster ids int32 [n], n_clusters, modularity of the best start)."""
    global _ref_lib
    if _ref_lib is None:
        path = build_ref()
        if path is None:
            raise RuntimeError("oracle/_ref/libmodopt_ref.so is not built and /root/reference is absent")
        L = C.CDLL(path)
        L.ref_modularity_cluster.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                             C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ref_modularity_cluster.restype = C.c_int
        _ref_lib = L
    p = np.ascontiguousarray(p, dtype=np.int32)
    i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = p.size - 1
    out = np.zeros(n, dtype=np.int32)
    q = C.c_double(0.0)
    rc = _ref_lib.ref_modularity_cluster(n, _p(p), _p(i), _p(x), int(modularity_function), float(resolution),
                                         int(algorithm), int(n_random_starts), int(n_iterations),
                                         int(random_seed), _p(out), C.addressof(q))
    if rc < 0:
        raise ValueError({-1: "invalid argument", -2: "Matrix contained no network data.  Check format.",
                          -3: "Clustering step failed.", -4: "C++ exception"}.get(rc, str(rc)))
    return out, rc, q.value


def modularity_cluster_file_ref(filename, modularity_function=1, resolution=0.8, algorithm=1, n_random_starts=10,
                                n_iterations=10, random_seed=0, capacity=1 << 22):
    """RunModularityClusteringCpp(edgefilename = filename) of the reference itself (its own readInputFile).
    Returns (cluster ids int32 [n_nodes], n_clusters, modularity)."""
    modularity_cluster_ref(np.array([0, 1, 2]), np.array([1, 0]), np.array([1.0, 1.0]), n_random_starts=1,
                           n_iterations=1)                    # loads the library
    L = _ref_lib
    L.ref_modularity_cluster_file.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                              C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_modularity_cluster_file.restype = C.c_int
    out = np.zeros(capacity, dtype=np.int32)
    n = C.c_int(0)
    q = C.c_double(0.0)
    rc = L.ref_modularity_cluster_file(str(filename).encode(), int(modularity_function), float(resolution),
                                       int(algorithm), int(n_random_starts), int(n_iterations), int(random_seed),
                                       _p(out), capacity, C.addressof(n), C.addressof(q))
    if rc < 0:
        raise ValueError({-1: "invalid argument", -3: "Clustering step failed.", -4: "C++ exception",
                          -5: "Could not parse edge file."}.get(rc, str(rc)))
    return out[:n.value].copy(), rc, q.value


# ----------------------------------------------------------------------------------------------
# SNN side consumers (src/snn.cpp:40-126, src/fast_NN_dist.cpp)
# ----------------------------------------------------------------------------------------------
def write_edge_file(p, i, x, filename: str) -> int:
    p = np.ascontiguousarray(p, dtype=np.int32)
    i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    return int(lib().oracle_write_edge_file(_p(p), _p(i), _p(x), p.size - 1, filename.encode()))


def fast_dist(x, y, lists):
    """lists: one 1-based index vector per row of x -> list of distance vectors (fast_dist)"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(v) for v in lists])
    nbr = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.int32) for v in lists]) if ptr[-1] else
                               np.zeros(0, np.int32), dtype=np.int32)
    out = np.empty(int(ptr[-1]), dtype=np.float64)
    if lib().oracle_fast_dist(_p(x), x.shape[0], _p(y), y.shape[0], x.shape[1], _p(ptr), _p(nbr), _p(out)) != 0:
        raise ValueError("neighbour index out of range")
    return [out[ptr[r]:ptr[r + 1]] for r in range(len(lists))]


def snn_smallest_nonzero_dist(p, i, x, mat, n_small: int, nearest_dist):
    p = np.ascontiguousarray(p, dtype=np.int32)
    i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    mat = np.ascontiguousarray(mat, dtype=np.float64)
    nearest = np.ascontiguousarray(nearest_dist, dtype=np.float64)
    n = p.size - 1
    out = np.empty(n, dtype=np.float64)
    lib().oracle_snn_smallest_nonzero_dist(_p(p), _p(i), _p(x), n, _p(mat), mat.shape[1], int(n_small), _p(nearest),
                                           _p(out))
    return out
