"""ctypes binding of libb200nn.so -- the only compute back end of this package.

There is deliberately no CPU fallback: if the CUDA library is missing or no
B200 is visible, every compute call raises :class:`B200Error`.
"""
from __future__ import annotations

import ctypes as C
import weakref
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200nn.so")

OK, ERR_INVALID, ERR_K_TOO_LARGE, ERR_CUDA, ERR_INDEX_RANGE, ERR_NNZ_OVERFLOW, ERR_NOMEM, ERR_STATE = range(8)
COL_MAJOR, ROW_MAJOR = 0, 1
KNN_AUTO, KNN_EXACT, KNN_TENSOR = 0, 1, 2

# every symbol include/b200nn.h declares (tests check the library exports them all)
SYMBOLS = [
    "b200nn_ctx_create", "b200nn_ctx_destroy", "b200nn_ctx_sync", "b200nn_ctx_stream",
    "b200nn_last_error", "b200nn_version", "b200nn_device_count",
    "b200nn_knn", "b200nn_snn_count", "b200nn_snn_fill",
    "b200nn_nn_graph_count", "b200nn_nn_graph_fill",
    "b200nn_find_neighbors_count", "b200nn_find_neighbors_fill",
    "b200nn_index_create", "b200nn_index_destroy", "b200nn_index_rows", "b200nn_index_dims", "b200nn_index_knn",
    "b200nn_dev_knn", "b200nn_dev_knn_candidates", "b200nn_dev_graphs", "b200nn_dev_result", "b200nn_dev_copy_result",
    "b200nn_multi_create", "b200nn_multi_destroy", "b200nn_multi_size", "b200nn_multi_knn",
    "b200nn_multi_find_neighbors_count", "b200nn_multi_find_neighbors_fill",
    "b200nn_modularity_cluster", "b200nn_modularity_cluster_resident",
    "b200nn_modularity_cluster_edges", "b200nn_modularity_cluster_edge_file",
    "b200nn_write_edge_file", "b200nn_fast_dist", "b200nn_snn_smallest_nonzero_dist",
    "b200nn_dev_embedding_prepare",
]


class B200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libb200nn error {code}: {message}")
        self.code = code
        self.message = message


class Opts(C.Structure):
    _fields_ = [("layout", C.c_int32), ("knn_algo", C.c_int32), ("n_cand", C.c_int32),
                ("scan_mode", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("ms_h2d", "ms_prep", "ms_knn_cand", "ms_knn_rerank", "ms_knn_fallback",
                 "ms_nn_graph", "ms_snn", "ms_d2h")] + \
               [(n, C.c_int64) for n in
                ("n_uncertified", "nnz_nn", "nnz_snn", "snn_expanded", "max_indegree",
                 "gpu_launches", "knn_tiles_visited", "knn_tiles_total")] + \
               [("ms_knn_pass2", C.c_double)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class ModoptStats(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("n_edges", C.c_int64), ("n_clusters", C.c_int32), ("pad_", C.c_int32),
                ("modularity", C.c_double)] + \
               [(n, C.c_int64) for n in ("windows", "rounds", "sequential_steps", "window_retries",
                                         "local_moving_calls", "levels", "max_window")] + \
               [(n, C.c_double) for n in ("ms_total", "ms_build", "ms_local_moving", "ms_reduce", "ms_quality", "ms_subnetworks")] + \
               [("gpu_launches", C.c_int64)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "pad_"}


class DevCsr(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("nnz", C.c_int64), ("p", C.c_void_p), ("i", C.c_void_p),
                ("x", C.c_void_p)]


_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libb200nn.so, failing loudly when it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200Error(ERR_CUDA, f"{LIB_PATH} is missing: build it with "
                                      "`python -m seurat_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.b200nn_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.b200nn_ctx_destroy.argtypes = [vp]
        L.b200nn_ctx_destroy.restype = None
        L.b200nn_ctx_sync.argtypes = [vp]
        L.b200nn_ctx_stream.argtypes = [vp]
        L.b200nn_ctx_stream.restype = vp
        L.b200nn_last_error.restype = C.c_char_p
        L.b200nn_knn.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp, vp, C.POINTER(Opts), C.POINTER(Stats)]
        L.b200nn_snn_count.argtypes = [vp, vp, i64, i32, dbl, C.POINTER(Opts), C.POINTER(i64), C.POINTER(Stats)]
        L.b200nn_snn_fill.argtypes = [vp, vp, vp, vp]
        L.b200nn_nn_graph_count.argtypes = [vp, vp, i64, i32, i64, C.POINTER(Opts), C.POINTER(i64), C.POINTER(Stats)]
        L.b200nn_nn_graph_fill.argtypes = [vp, vp, vp, vp]
        L.b200nn_find_neighbors_count.argtypes = [vp, vp, i64, i32, i32, dbl, i32, vp, vp, vp, vp, vp,
                                                  C.POINTER(Opts), C.POINTER(i64), C.POINTER(i64), C.POINTER(Stats)]
        L.b200nn_find_neighbors_fill.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b200nn_index_create.argtypes = [vp, vp, i64, i32, C.POINTER(Opts), C.POINTER(vp)]
        L.b200nn_index_destroy.argtypes = [vp, vp]
        L.b200nn_index_destroy.restype = None
        L.b200nn_index_rows.argtypes = [vp]
        L.b200nn_index_rows.restype = i64
        L.b200nn_index_dims.argtypes = [vp]
        L.b200nn_index_dims.restype = i32
        L.b200nn_index_knn.argtypes = [vp, vp, vp, i64, i32, vp, vp, C.POINTER(Opts), C.POINTER(Stats)]
        L.b200nn_dev_knn.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp, vp, C.POINTER(Opts), C.POINTER(Stats)]
        L.b200nn_dev_knn_candidates.argtypes = [vp, vp, i64, vp, i64, i32, i32, i32, vp, vp, vp, vp]
        L.b200nn_dev_graphs.argtypes = [vp, vp, i64, i32, dbl, i32, i64, i64, C.POINTER(i64), C.POINTER(i64),
                                        C.POINTER(Stats)]
        L.b200nn_dev_result.argtypes = [vp, i32, C.POINTER(DevCsr)]
        L.b200nn_dev_copy_result.argtypes = [vp, i32, vp, vp, vp]
        L.b200nn_multi_create.argtypes = [vp, i32, C.POINTER(vp)]
        L.b200nn_multi_destroy.argtypes = [vp]
        L.b200nn_multi_destroy.restype = None
        L.b200nn_multi_size.argtypes = [vp]
        L.b200nn_multi_size.restype = i32
        L.b200nn_multi_knn.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp, vp, C.POINTER(Opts), C.POINTER(Stats)]
        L.b200nn_multi_find_neighbors_count.argtypes = [vp, vp, i64, i32, i32, dbl, i32, vp, vp, vp, vp, vp,
                                                        C.POINTER(Opts), C.POINTER(i64), C.POINTER(i64),
                                                        C.POINTER(Stats)]
        L.b200nn_multi_find_neighbors_fill.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b200nn_modularity_cluster.argtypes = [vp, i64, vp, vp, vp, i32, dbl, i32, i32, i32, C.c_uint64, vp,
                                                C.POINTER(ModoptStats)]
        L.b200nn_modularity_cluster_resident.argtypes = [vp, i32, dbl, i32, i32, i32, C.c_uint64, vp, i64,
                                                         C.POINTER(ModoptStats)]
        L.b200nn_modularity_cluster_edges.argtypes = [vp, i64, vp, vp, vp, i64, i32, dbl, i32, i32, i32, C.c_uint64, vp,
                                                      C.POINTER(ModoptStats)]
        L.b200nn_modularity_cluster_edge_file.argtypes = [vp, C.c_char_p, i32, dbl, i32, i32, i32, C.c_uint64, vp, i64,
                                                          C.POINTER(i64), C.POINTER(ModoptStats)]
        L.b200nn_dev_embedding_prepare.argtypes = [vp, vp, i64, i32, i32, vp, i32, i32, vp]
        L.b200nn_write_edge_file.argtypes = [vp, vp, vp, i64, C.c_char_p, C.POINTER(i64)]
        L.b200nn_fast_dist.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp, vp, vp]
        L.b200nn_snn_smallest_nonzero_dist.argtypes = [vp, vp, vp, vp, i64, vp, i32, i32, i32, vp, vp]
        _lib = L
        return L


def check(rc: int) -> None:
    if rc != OK:
        raise B200Error(rc, load().b200nn_last_error().decode("utf-8", "replace"))


def write_edge_file(p, i, x, filename: str) -> int:
    """WriteEdgeFile (src/snn.cpp:40-59): strict lower triangle as "col\\trow\\tvalue" lines; host I/O,
    needs no device.  Returns the number of lines written."""
    p = np.ascontiguousarray(p, dtype=np.int32)
    i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n_lines = C.c_int64(0)
    check(load().b200nn_write_edge_file(_ptr(p), _ptr(i), _ptr(x), p.size - 1, str(filename).encode(),
                                        C.byref(n_lines)))
    return int(n_lines.value)


def device_count() -> int:
    return int(load().b200nn_device_count())


def _ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())  # torch tensor


class Context:
    """One library context = one CUDA device, one stream, one reusable workspace."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p(None)
        self.device = device
        L = load()
        check(L.b200nn_ctx_create(int(device), C.byref(self._h)))
        self._L = L
        self._indexes = weakref.WeakSet()   # resident reference sets; they must go before the context

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            for ix in list(getattr(self, "_indexes", ())):
                ix.close()
            self._L.b200nn_ctx_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self) -> None:
        check(self._L.b200nn_ctx_sync(self._h))

    @property
    def stream_ptr(self) -> int:
        return int(self._L.b200nn_ctx_stream(self._h) or 0)

    # ---- modularity optimiser (the step after FindNeighbors) -------------------
    def modularity_cluster(self, p, i, x, modularity_function: int = 1, resolution: float = 0.8,
                           algorithm: int = 1, n_random_starts: int = 10, n_iterations: int = 10,
                           random_seed: int = 0):
        """dgCMatrix slots (p, i, x) of an n x n graph in host memory -> (cluster ids int32 [n], stats)."""
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = p.size - 1
        out = np.empty(max(n, 0), dtype=np.int32)
        st = ModoptStats()
        check(self._L.b200nn_modularity_cluster(self._h, n, _ptr(p), _ptr(i), _ptr(x), int(modularity_function),
                                                float(resolution), int(algorithm), int(n_random_starts),
                                                int(n_iterations), int(random_seed) & (2 ** 64 - 1), _ptr(out),
                                                C.byref(st)))
        return out, st.as_dict()

    def modularity_cluster_resident(self, n: int, modularity_function: int = 1, resolution: float = 0.8,
                                    algorithm: int = 1, n_random_starts: int = 10, n_iterations: int = 10,
                                    random_seed: int = 0):
        """The same on the SNN graph still resident in HBM after find_neighbors / dev_graphs."""
        out = np.empty(n, dtype=np.int32)
        st = ModoptStats()
        check(self._L.b200nn_modularity_cluster_resident(self._h, int(modularity_function), float(resolution),
                                                         int(algorithm), int(n_random_starts), int(n_iterations),
                                                         int(random_seed) & (2 ** 64 - 1), _ptr(out), n,
                                                         C.byref(st)))
        return out, st.as_dict()

    def modularity_cluster_edges(self, node1, node2, weight, n_nodes: int, modularity_function: int = 1,
                                 resolution: float = 0.8, algorithm: int = 1, n_random_starts: int = 10,
                                 n_iterations: int = 10, random_seed: int = 0):
        """edge list (0-based node ids, weight may be None = 1) -> (cluster ids int32 [n_nodes], stats)"""
        node1 = np.ascontiguousarray(node1, dtype=np.int32)
        node2 = np.ascontiguousarray(node2, dtype=np.int32)
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
        out = np.empty(max(int(n_nodes), 0), dtype=np.int32)
        st = ModoptStats()
        check(self._L.b200nn_modularity_cluster_edges(self._h, int(n_nodes), _ptr(node1), _ptr(node2), _ptr(w),
                                                      node1.size, int(modularity_function), float(resolution),
                                                      int(algorithm), int(n_random_starts), int(n_iterations),
                                                      int(random_seed) & (2 ** 64 - 1), _ptr(out), C.byref(st)))
        return out, st.as_dict()

    def modularity_cluster_edge_file(self, filename: str, modularity_function: int = 1, resolution: float = 0.8,
                                     algorithm: int = 1, n_random_starts: int = 10, n_iterations: int = 10,
                                     random_seed: int = 0):
        """the edge.file.name input of RunModularityClusteringCpp: "node1<TAB>node2[<TAB>weight]" lines"""
        n = C.c_int64(0)
        args = (int(modularity_function), float(resolution), int(algorithm), int(n_random_starts), int(n_iterations),
                int(random_seed) & (2 ** 64 - 1))
        fn = str(filename).encode()
        check(self._L.b200nn_modularity_cluster_edge_file(self._h, fn, *args, None, 0, C.byref(n), None))
        out = np.empty(n.value, dtype=np.int32)
        st = ModoptStats()
        check(self._L.b200nn_modularity_cluster_edge_file(self._h, fn, *args, _ptr(out), n.value, C.byref(n),
                                                          C.byref(st)))
        return out, st.as_dict()

    # ---- SNN side consumers (src/snn.cpp:40-126, src/fast_NN_dist.cpp) ---------------
    def fast_dist(self, x, y, ptr, nbr, layout: int):
        """distances between row r of x and the rows of y listed (1-based) in nbr[ptr[r]:ptr[r+1]]"""
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        nbr = np.ascontiguousarray(nbr, dtype=np.int32)
        out = np.empty(int(ptr[-1]), dtype=np.float64)
        check(self._L.b200nn_fast_dist(self._h, _ptr(x), x.shape[0], _ptr(y), y.shape[0], x.shape[1], layout,
                                       _ptr(ptr), _ptr(nbr), _ptr(out)))
        return out

    def snn_smallest_nonzero_dist(self, p, i, x, mat, layout: int, n_small: int, nearest_dist):
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        nearest = np.ascontiguousarray(nearest_dist, dtype=np.float64)
        n = p.size - 1
        out = np.empty(n, dtype=np.float64)
        check(self._L.b200nn_snn_smallest_nonzero_dist(self._h, _ptr(p), _ptr(i), _ptr(x), n, _ptr(mat), mat.shape[1],
                                                       layout, int(n_small), _ptr(nearest), _ptr(out)))
        return out

    def index(self, ref, layout: int) -> "Index":
        """Keep the reference embedding `ref` [nr, d] on the device for repeated queries."""
        return Index(self, ref, layout)

    # ---- host-buffer calls ------------------------------------------------
    def knn(self, ref, qry, k: int, layout: int, algo: int = KNN_AUTO, n_cand: int = 0, scan_mode: int = 0):
        """ref/qry: float64 arrays [n, d] stored in `layout`.  Returns (idx, dist, stats)."""
        nr, d = ref.shape
        nq = nr if qry is None else qry.shape[0]
        order = "C" if layout == ROW_MAJOR else "F"
        idx = np.empty((nq, k), dtype=np.int32, order=order)
        dist = np.empty((nq, k), dtype=np.float64, order=order)
        o = Opts(layout, algo, n_cand, scan_mode)
        st = Stats()
        check(self._L.b200nn_knn(self._h, _ptr(ref), nr, _ptr(qry), nq, d, k, _ptr(idx), _ptr(dist),
                                 C.byref(o), C.byref(st)))
        return idx, dist, st.as_dict()

    def compute_snn(self, nn_ranked, prune: float, layout: int):
        n, k = nn_ranked.shape
        o = Opts(layout, 0, 0, 0)
        st = Stats()
        nnz = C.c_int64(0)
        check(self._L.b200nn_snn_count(self._h, _ptr(nn_ranked), n, k, float(prune), C.byref(o),
                                       C.byref(nnz), C.byref(st)))
        p = np.empty(n + 1, dtype=np.int32)
        i = np.empty(nnz.value, dtype=np.int32)
        x = np.empty(nnz.value, dtype=np.float64)
        check(self._L.b200nn_snn_fill(self._h, _ptr(p), _ptr(i), _ptr(x)))
        return p, i, x, st.as_dict()

    def nn_graph(self, nn_ranked, n: int, layout: int):
        rows, k = nn_ranked.shape
        o = Opts(layout, 0, 0, 0)
        st = Stats()
        nnz = C.c_int64(0)
        check(self._L.b200nn_nn_graph_count(self._h, _ptr(nn_ranked), rows, k, n, C.byref(o),
                                            C.byref(nnz), C.byref(st)))
        p = np.empty(n + 1, dtype=np.int32)
        i = np.empty(nnz.value, dtype=np.int32)
        x = np.empty(nnz.value, dtype=np.float64)
        check(self._L.b200nn_nn_graph_fill(self._h, _ptr(p), _ptr(i), _ptr(x)))
        return p, i, x, st.as_dict()

    def find_neighbors(self, data, k: int, prune: float, compute_snn: bool, layout: int,
                       algo: int = KNN_AUTO, n_cand: int = 0, want_knn: bool = True,
                       out: dict | None = None, scan_mode: int = 0):
        """Fused kNN + nn Graph (+ SNN) on host buffers.  `out` may carry
        preallocated (pinned) result arrays keyed idx/dist/nn_p/nn_i/nn_x/snn_p/snn_i/snn_x."""
        n, d = data.shape
        order = "C" if layout == ROW_MAJOR else "F"
        out = out or {}
        idx = dist = None
        if want_knn:
            idx = out.get("idx")
            dist = out.get("dist")
            if idx is None:
                idx = np.empty((n, k), dtype=np.int32, order=order)
            if dist is None:
                dist = np.empty((n, k), dtype=np.float64, order=order)
        o = Opts(layout, algo, n_cand, scan_mode)
        st = Stats()
        nnz_nn, nnz_snn = C.c_int64(0), C.c_int64(0)
        # nn Graph buffers handed in by the caller (capacity n * k) are filled during the count call,
        # while the SNN stage runs
        early = {}
        for name, count, dt in (("nn_p", n + 1, np.int32), ("nn_i", n * k, np.int32), ("nn_x", n * k, np.float64)):
            a = out.get(name)
            if a is not None and a.size >= count and a.dtype == dt and a.flags.c_contiguous:
                early[name] = a
        if len(early) != 3:
            early = {}
        check(self._L.b200nn_find_neighbors_count(self._h, _ptr(data), n, d, k, float(prune),
                                                  1 if compute_snn else 0, _ptr(idx), _ptr(dist),
                                                  _ptr(early.get("nn_p")), _ptr(early.get("nn_i")),
                                                  _ptr(early.get("nn_x")),
                                                  C.byref(o), C.byref(nnz_nn), C.byref(nnz_snn),
                                                  C.byref(st)))

        def buf(name, count, dtype):
            a = out.get(name)
            if a is not None and a.size >= count:
                return a[:count] if a.size != count else a
            return np.empty(count, dtype=dtype)

        if early:
            nn_p, nn_i, nn_x = early["nn_p"][:n + 1], early["nn_i"][:nnz_nn.value], early["nn_x"][:nnz_nn.value]
            fill_nn = (None, None, None)
        else:
            nn_p, nn_i, nn_x = buf("nn_p", n + 1, np.int32), buf("nn_i", nnz_nn.value, np.int32), \
                buf("nn_x", nnz_nn.value, np.float64)
            fill_nn = (nn_p, nn_i, nn_x)
        snn = (None, None, None)
        if compute_snn:
            snn = (buf("snn_p", n + 1, np.int32), buf("snn_i", nnz_snn.value, np.int32),
                   buf("snn_x", nnz_snn.value, np.float64))
        if fill_nn[0] is not None or compute_snn:
            check(self._L.b200nn_find_neighbors_fill(self._h, _ptr(fill_nn[0]), _ptr(fill_nn[1]), _ptr(fill_nn[2]),
                                                     _ptr(snn[0]), _ptr(snn[1]), _ptr(snn[2])))
        return {"idx": idx, "dist": dist, "nn": (nn_p, nn_i, nn_x), "snn": snn if compute_snn else None,
                "stats": st.as_dict()}

    # ---- device-buffer calls (pointers or torch CUDA tensors) --------------
    def dev_knn(self, d_ref, nr: int, d_qry, nq: int, d: int, k: int, d_idx0, d_dist,
                algo: int = KNN_AUTO, n_cand: int = 0, scan_mode: int = 0, shard_rank: int = 0,
                shard_world: int = 0) -> dict:
        o = Opts(ROW_MAJOR, algo, n_cand, scan_mode, shard_rank, shard_world)
        st = Stats()
        check(self._L.b200nn_dev_knn(self._h, _ptr(d_ref), nr, _ptr(d_qry), nq, d, k, _ptr(d_idx0),
                                     _ptr(d_dist), C.byref(o), C.byref(st)))
        return st.as_dict()

    def dev_knn_candidates(self, d_ref, nr: int, d_qry, nq: int, d: int, n_cand: int, d_cand, d_tau,
                           d_keys, scale_mu: np.ndarray | None = None, scan_mode: int = 0) -> None:
        check(self._L.b200nn_dev_knn_candidates(self._h, _ptr(d_ref), nr, _ptr(d_qry), nq, d, n_cand,
                                                scan_mode, _ptr(d_cand), _ptr(d_tau), _ptr(d_keys),
                                                _ptr(scale_mu)))

    def dev_embedding_prepare(self, d_src, n: int, n_cols: int, layout: int, dims, l2_norm: bool, d_out) -> None:
        """columns `dims` (0-based, None = all) of a device matrix [n x n_cols] in `layout`, optionally L2
        normalised like R's L2Norm, as the row-major device matrix the dev_* stages take"""
        dd = None if dims is None else np.ascontiguousarray(dims, dtype=np.int32)
        check(self._L.b200nn_dev_embedding_prepare(self._h, _ptr(d_src), n, n_cols, layout, _ptr(dd),
                                                   0 if dd is None else dd.size, 1 if l2_norm else 0, _ptr(d_out)))

    def dev_graphs(self, d_idx0, n: int, k: int, prune: float, compute_snn: bool, row_begin: int,
                   row_end: int) -> dict:
        st = Stats()
        nnz_nn, nnz_snn = C.c_int64(0), C.c_int64(0)
        check(self._L.b200nn_dev_graphs(self._h, _ptr(d_idx0), n, k, float(prune),
                                        1 if compute_snn else 0, row_begin, row_end,
                                        C.byref(nnz_nn), C.byref(nnz_snn), C.byref(st)))
        r = st.as_dict()
        r["nnz_nn"], r["nnz_snn"] = nnz_nn.value, nnz_snn.value
        return r

    def dev_result(self, which: int) -> DevCsr:
        r = DevCsr()
        check(self._L.b200nn_dev_result(self._h, which, C.byref(r)))
        return r

    def dev_copy_result(self, which: int, d_p, d_i, d_x) -> None:
        """Device-to-device copy of a result into caller buffers (p is int64)."""
        check(self._L.b200nn_dev_copy_result(self._h, which, _ptr(d_p), _ptr(d_i), _ptr(d_x)))


class MultiContext:
    """Several devices driven from this process (b200nn_multi_*): what the R shim creates when more
    than one GPU is visible.  ``knn`` and ``find_neighbors`` have the signatures of :class:`Context`,
    so ``FindNeighbors(..., ctx=MultiContext([0, 1, 2, 3]))`` shards the call; the stages that take an
    existing index (``ComputeSNN`` / nn Graph alone, resident indexes) run on the first device."""

    def __init__(self, devices=None):
        L = load()
        self._L = L
        self._h = C.c_void_p(None)
        self._single = None
        if devices is None:
            arr, n = None, 0
        else:
            devices = [int(x) for x in devices]
            arr, n = (C.c_int32 * len(devices))(*devices), len(devices)
        check(L.b200nn_multi_create(arr, n, C.byref(self._h)))
        self.devices = devices if devices is not None else list(range(self.size))

    @property
    def size(self) -> int:
        return int(self._L.b200nn_multi_size(self._h))

    @property
    def device(self) -> int:
        return self.devices[0]

    def _first(self) -> Context:
        if self._single is None:
            self._single = Context(self.devices[0])
        return self._single

    def close(self) -> None:
        if getattr(self, "_single", None) is not None:
            self._single.close()
            self._single = None
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.b200nn_multi_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def knn(self, ref, qry, k: int, layout: int, algo: int = KNN_AUTO, n_cand: int = 0, scan_mode: int = 0):
        nr, d = ref.shape
        nq = nr if qry is None else qry.shape[0]
        order = "C" if layout == ROW_MAJOR else "F"
        idx = np.empty((nq, k), dtype=np.int32, order=order)
        dist = np.empty((nq, k), dtype=np.float64, order=order)
        o = Opts(layout, algo, n_cand, scan_mode)
        st = Stats()
        check(self._L.b200nn_multi_knn(self._h, _ptr(ref), nr, _ptr(qry), nq, d, k, _ptr(idx), _ptr(dist),
                                       C.byref(o), C.byref(st)))
        return idx, dist, st.as_dict()

    def find_neighbors(self, data, k: int, prune: float, compute_snn: bool, layout: int,
                       algo: int = KNN_AUTO, n_cand: int = 0, want_knn: bool = True,
                       out: dict | None = None, scan_mode: int = 0, pinned_snn: bool = False):
        """Same contract as :meth:`Context.find_neighbors`.  ``pinned_snn`` allocates the SNN arrays
        page-locked (through torch) when the caller did not pass ``snn_*`` buffers."""
        n, d = data.shape
        order = "C" if layout == ROW_MAJOR else "F"
        out = out or {}
        idx = dist = None
        if want_knn:
            idx, dist = out.get("idx"), out.get("dist")
            if idx is None:
                idx = np.empty((n, k), dtype=np.int32, order=order)
            if dist is None:
                dist = np.empty((n, k), dtype=np.float64, order=order)
        o = Opts(layout, algo, n_cand, scan_mode)
        st = Stats()
        nnz_nn, nnz_snn = C.c_int64(0), C.c_int64(0)
        early = {}
        for name, count, dt in (("nn_p", n + 1, np.int32), ("nn_i", n * k, np.int32), ("nn_x", n * k, np.float64)):
            a = out.get(name)
            if a is not None and a.size >= count and a.dtype == dt and a.flags.c_contiguous:
                early[name] = a
        if len(early) != 3:
            early = {}
        check(self._L.b200nn_multi_find_neighbors_count(self._h, _ptr(data), n, d, k, float(prune),
                                                        1 if compute_snn else 0, _ptr(idx), _ptr(dist),
                                                        _ptr(early.get("nn_p")), _ptr(early.get("nn_i")),
                                                        _ptr(early.get("nn_x")), C.byref(o), C.byref(nnz_nn),
                                                        C.byref(nnz_snn), C.byref(st)))

        def buf(name, count, dtype, pinned=False):
            a = out.get(name)
            if a is not None and a.size >= count:
                return a[:count] if a.size != count else a
            if pinned:
                import torch
                return torch.empty(count, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory().numpy()
            return np.empty(count, dtype=dtype)

        if early:
            nn_p, nn_i, nn_x = early["nn_p"][:n + 1], early["nn_i"][:nnz_nn.value], early["nn_x"][:nnz_nn.value]
            fill_nn = (None, None, None)
        else:
            nn_p, nn_i, nn_x = buf("nn_p", n + 1, np.int32), buf("nn_i", nnz_nn.value, np.int32), \
                buf("nn_x", nnz_nn.value, np.float64)
            fill_nn = (nn_p, nn_i, nn_x)
        snn = (None, None, None)
        if compute_snn:
            snn = (buf("snn_p", n + 1, np.int32, pinned_snn), buf("snn_i", nnz_snn.value, np.int32, pinned_snn),
                   buf("snn_x", nnz_snn.value, np.float64, pinned_snn))
        if fill_nn[0] is not None or compute_snn:
            check(self._L.b200nn_multi_find_neighbors_fill(self._h, _ptr(fill_nn[0]), _ptr(fill_nn[1]),
                                                           _ptr(fill_nn[2]), _ptr(snn[0]), _ptr(snn[1]),
                                                           _ptr(snn[2])))
        return {"idx": idx, "dist": dist, "nn": (nn_p, nn_i, nn_x), "snn": snn if compute_snn else None,
                "stats": st.as_dict()}

    # stages that start from an existing index run on the first device
    def compute_snn(self, nn_ranked, prune: float, layout: int):
        return self._first().compute_snn(nn_ranked, prune, layout)

    def nn_graph(self, nn_ranked, n: int, layout: int):
        return self._first().nn_graph(nn_ranked, n, layout)

    # ---- modularity optimiser (the step after FindNeighbors) -------------------
    def modularity_cluster(self, p, i, x, modularity_function: int = 1, resolution: float = 0.8,
                           algorithm: int = 1, n_random_starts: int = 10, n_iterations: int = 10,
                           random_seed: int = 0):
        """dgCMatrix slots (p, i, x) of an n x n graph in host memory -> (cluster ids int32 [n], stats)."""
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = p.size - 1
        out = np.empty(max(n, 0), dtype=np.int32)
        st = ModoptStats()
        check(self._L.b200nn_modularity_cluster(self._h, n, _ptr(p), _ptr(i), _ptr(x), int(modularity_function),
                                                float(resolution), int(algorithm), int(n_random_starts),
                                                int(n_iterations), int(random_seed) & (2 ** 64 - 1), _ptr(out),
                                                C.byref(st)))
        return out, st.as_dict()

    def modularity_cluster_resident(self, n: int, modularity_function: int = 1, resolution: float = 0.8,
                                    algorithm: int = 1, n_random_starts: int = 10, n_iterations: int = 10,
                                    random_seed: int = 0):
        """The same on the SNN graph still resident in HBM after find_neighbors / dev_graphs."""
        out = np.empty(n, dtype=np.int32)
        st = ModoptStats()
        check(self._L.b200nn_modularity_cluster_resident(self._h, int(modularity_function), float(resolution),
                                                         int(algorithm), int(n_random_starts), int(n_iterations),
                                                         int(random_seed) & (2 ** 64 - 1), _ptr(out), n,
                                                         C.byref(st)))
        return out, st.as_dict()

    # ---- SNN side consumers (src/snn.cpp:40-126, src/fast_NN_dist.cpp) ---------------
    def fast_dist(self, x, y, ptr, nbr, layout: int):
        """distances between row r of x and the rows of y listed (1-based) in nbr[ptr[r]:ptr[r+1]]"""
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        nbr = np.ascontiguousarray(nbr, dtype=np.int32)
        out = np.empty(int(ptr[-1]), dtype=np.float64)
        check(self._L.b200nn_fast_dist(self._h, _ptr(x), x.shape[0], _ptr(y), y.shape[0], x.shape[1], layout,
                                       _ptr(ptr), _ptr(nbr), _ptr(out)))
        return out

    def snn_smallest_nonzero_dist(self, p, i, x, mat, layout: int, n_small: int, nearest_dist):
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        nearest = np.ascontiguousarray(nearest_dist, dtype=np.float64)
        n = p.size - 1
        out = np.empty(n, dtype=np.float64)
        check(self._L.b200nn_snn_smallest_nonzero_dist(self._h, _ptr(p), _ptr(i), _ptr(x), n, _ptr(mat), mat.shape[1],
                                                       layout, int(n_small), _ptr(nearest), _ptr(out)))
        return out

    def index(self, ref, layout: int) -> "Index":
        return self._first().index(ref, layout)


_default_ctx: dict[int, Context] = {}


class Index:
    """Device-resident reference set (b200nn_index): what NNHelper(cache.index = TRUE) returns as
    `alg.idx` and FindNeighbors(index = ) accepts -- R/clustering.R:596-597, :1686-1688."""

    def __init__(self, ctx: Context, ref, layout: int):
        ref = np.asarray(ref)
        if ref.dtype != np.float64 or ref.ndim != 2:
            raise TypeError("the reference embedding must be a float64 matrix")
        self._ctx = ctx
        self._h = C.c_void_p(None)
        o = Opts(layout, KNN_AUTO, 0, 0)
        check(ctx._L.b200nn_index_create(ctx._h, _ptr(ref), ref.shape[0], ref.shape[1], C.byref(o),
                                         C.byref(self._h)))
        self.layout = layout
        ctx._indexes.add(self)

    @property
    def rows(self) -> int:
        return int(self._ctx._L.b200nn_index_rows(self._h))

    @property
    def dims(self) -> int:
        return int(self._ctx._L.b200nn_index_dims(self._h))

    def knn(self, qry, k: int, algo: int = KNN_AUTO, n_cand: int = 0, scan_mode: int = 0):
        """qry: float64 [nq, d] in the index's layout, or None for the self query."""
        if not self._h.value:
            raise B200Error(7, "the index was destroyed")
        nq = self.rows if qry is None else qry.shape[0]
        order = "C" if self.layout == ROW_MAJOR else "F"
        idx = np.empty((nq, k), dtype=np.int32, order=order)
        dist = np.empty((nq, k), dtype=np.float64, order=order)
        o = Opts(self.layout, algo, n_cand, scan_mode)
        st = Stats()
        check(self._ctx._L.b200nn_index_knn(self._ctx._h, self._h, _ptr(qry), nq, k, _ptr(idx), _ptr(dist),
                                            C.byref(o), C.byref(st)))
        return idx, dist, st.as_dict()

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value and self._ctx._h.value:
            self._ctx._L.b200nn_index_destroy(self._ctx._h, self._h)
        self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
