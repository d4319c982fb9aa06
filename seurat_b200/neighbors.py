"""Host-side mirror of Seurat's FindNeighbors interface over libb200nn.

R is not available in the build image, so the host layer that the R package
would keep (argument handling in ``FindNeighbors.default``, ``NNHelper``, the
``Neighbor`` / ``Graph`` containers) is mirrored here in Python with the same
names, argument meaning, warnings and errors; the R-side `.Call` shim that
binds the same C ABI is in ``shim/`` and described in ``INTEGRATION.md``.

Reference (paths relative to the Seurat tree):
  FindNeighbors.default   R/clustering.R:582-685
  FindNeighbors.Seurat    R/clustering.R:790-898   (``FindNeighborsSeurat`` below)
  NNHelper                R/clustering.R:1658-1690
  ComputeSNN              src/snn.cpp:14-38 via R/RcppExports.R:96-98
  L2Norm                  R/utilities.R:2107-2122

All compute (kNN, nn Graph, SNN) runs in the CUDA library; there is no CPU path.
"""
from __future__ import annotations

import sys
import warnings
from dataclasses import dataclass, field

import numpy as np

from . import _lib

__all__ = ["WriteEdgeFile", "DirectSNNToFile", "fast_dist", "SNN_SmallestNonzero_Dist",
           "Embedding", "Neighbor", "Graph", "L2Norm", "NNHelper", "ComputeSNN", "FindNeighbors",
           "FindNeighborsSeurat", "SeuratLite", "Indices", "Distances", "Cells"]


def _message(msg: str, verbose: bool = True) -> None:
    if verbose:
        print(msg, file=sys.stderr)


# ---------------------------------------------------------------------------
# containers
# ---------------------------------------------------------------------------
class Embedding:
    """A numeric matrix with row names -- what R hands to ``FindNeighbors.default``.

    ``values`` is kept as float64; Fortran (column-major, R's layout) and C
    order are both accepted and passed to the library without a host copy.
    """

    def __init__(self, values, rownames=None):
        if hasattr(values, "index") and hasattr(values, "to_numpy"):  # pandas DataFrame
            if rownames is None:
                rownames = [str(x) for x in values.index]
            values = values.to_numpy()
        v = np.asarray(values, dtype=np.float64)
        if v.ndim == 2 and not (v.flags.c_contiguous or v.flags.f_contiguous):
            v = np.ascontiguousarray(v)
        self.values = v
        self.rownames = None if rownames is None else list(rownames)
        if self.rownames is not None and v.ndim == 2 and len(self.rownames) != v.shape[0]:
            raise ValueError("length of rownames must equal the number of rows")

    @property
    def shape(self):
        return self.values.shape

    def take_columns(self, dims) -> "Embedding":
        return Embedding(np.ascontiguousarray(self.values[:, list(dims)]), self.rownames)


def _as_embedding(x, what="object") -> Embedding:
    if isinstance(x, Embedding):
        return x
    return Embedding(x, getattr(x, "rownames", None))


def _layout_of(a: np.ndarray) -> int:
    # a C-contiguous [n, d] array is row-major; an F-contiguous one is R's layout
    return _lib.ROW_MAJOR if a.flags.c_contiguous else _lib.COL_MAJOR


@dataclass
class Neighbor:
    """SeuratObject's ``Neighbor`` class (slots nn.idx, nn.dist, alg.info, cell.names)."""
    nn_idx: np.ndarray            # int32 [n_query, k], 1-based rows of the reference data
    nn_dist: np.ndarray           # float64 [n_query, k]
    alg_info: dict = field(default_factory=dict)
    cell_names: list | None = None
    alg_idx: object = None


def Indices(object: Neighbor) -> np.ndarray:
    return object.nn_idx


def Distances(object: Neighbor) -> np.ndarray:
    return object.nn_dist


def Cells(object: Neighbor):
    return object.cell_names


class Graph:
    """SeuratObject's ``Graph`` (a dgCMatrix with an ``assay.used`` slot).

    Slots follow Matrix's dgCMatrix: ``p`` int32[ncol+1], ``i`` int32[nnz]
    (0-based, ascending inside each column), ``x`` float64[nnz], ``Dim``,
    ``Dimnames``.
    """

    def __init__(self, p, i, x, dim, dimnames=(None, None), assay_used=None):
        self.p, self.i, self.x = p, i, x
        self.Dim = (int(dim[0]), int(dim[1]))
        self.Dimnames = [dimnames[0], dimnames[1]]
        self.assay_used = assay_used

    @property
    def shape(self):
        return self.Dim

    @property
    def nnz(self) -> int:
        return int(self.i.shape[0])

    @property
    def rownames(self):
        return self.Dimnames[0]

    @rownames.setter
    def rownames(self, v):
        self.Dimnames[0] = None if v is None else list(v)

    @property
    def colnames(self):
        return self.Dimnames[1]

    @colnames.setter
    def colnames(self, v):
        self.Dimnames[1] = None if v is None else list(v)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.x, self.i, self.p), shape=self.Dim)


def as_Graph(x: Graph) -> Graph:
    return x


# ---------------------------------------------------------------------------
# helpers mirrored from R
# ---------------------------------------------------------------------------
def L2Norm(mat, MARGIN: int = 1):
    """Row (MARGIN=1) or column (2) L2 normalisation; non-finite results become 0.

    R/utilities.R:2107-2122.  R's ``sum`` accumulates in long double, which is
    reproduced here; this runs on the host exactly as in the reference (it is
    argument preparation, not part of the accelerated path).
    """
    emb = mat if isinstance(mat, Embedding) else None
    m = np.asarray(emb.values if emb is not None else mat, dtype=np.float64)
    axis = 1 if MARGIN == 1 else 0
    # sum(x^2) term by term in long double, in the order R visits them (numpy's own sum() is pairwise)
    sq = m * m
    acc = np.zeros(m.shape[1 - axis], dtype=np.longdouble)
    for j in range(m.shape[axis]):
        acc += sq[:, j] if axis == 1 else sq[j, :]
    norm = np.sqrt(acc.astype(np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = m / (norm[:, None] if axis == 1 else norm[None, :])
    out[~np.isfinite(out)] = 0.0
    return Embedding(out, emb.rownames) if emb is not None else out


def _check_finite(a: np.ndarray, what: str) -> None:
    # nn2 goes through .C with NAOK = FALSE: R refuses NA/NaN/Inf before the search starts
    if not np.isfinite(a).all():
        raise ValueError(f"NA/NaN/Inf in foreign function call (arg {what})")


def _warn_slow_path(st: dict, nq: int) -> None:
    """The result is exact either way; these inputs only make the search slow: rows whose
    certificate failed are recomputed by the FP64 scan, and an input without cluster structure (or one far
    outlier that dictates the FP16 scale) defeats the tile pruning."""
    if nq >= 10_000 and st.get("n_uncertified", 0) > 0.02 * nq:
        warnings.warn(f"{st['n_uncertified']} of {nq} query rows needed the FP64 fallback scan (exact duplicates, "
                      "lattice ties or badly scaled coordinates, e.g. a far outlier); the result is exact but slow",
                      RuntimeWarning, stacklevel=3)
    tot = st.get("knn_tiles_total", 0)
    if nq >= 200_000 and tot and st.get("knn_tiles_visited", 0) > 0.5 * tot:
        warnings.warn("the embedding has no cluster structure the tile pruning could use: the tensor-core search "
                      "scans nearly every tile (exact, but several times slower than on clustered data)",
                      RuntimeWarning, stacklevel=3)


def _nn2(data: np.ndarray, query: np.ndarray | None, k: int, ctx: _lib.Context, algo: int,
         n_cand: int, stats_out: dict | None, cache_index: bool = False, index: "_lib.Index | None" = None):
    """Drop-in for ``RANN::nn2(data, query, k, searchtype="standard", eps=0)``.

    ``index`` (a resident reference set from an earlier ``cache_index=True`` call) replaces
    ``data``; with ``cache_index`` the reference set is uploaded once, kept on the device and
    returned under ``"idx"`` -- the exact-search counterpart of the annoy index the reference
    caches (R/clustering.R:1686-1688)."""
    nr = index.rows if index is not None else data.shape[0]
    nd = index.dims if index is not None else data.shape[1]
    if k > nr:
        raise ValueError("Cannot find more nearest neighbours than there are points")
    if index is None:
        _check_finite(data, "data")
    if query is not None:
        _check_finite(query, "query")
        if query.shape[1] != nd:
            raise ValueError("Query and data tables must have same dimensions")
        want_c = (index.layout == _lib.ROW_MAJOR) if index is not None else data.flags.c_contiguous
        if query.flags.c_contiguous != want_c or not (query.flags.c_contiguous or query.flags.f_contiguous):
            query = np.asarray(query, order="C" if want_c else "F")
    if index is None and cache_index:
        index = ctx.index(data, _layout_of(data))
    if index is not None:
        idx, dist, st = index.knn(query, int(k), algo, n_cand)
    else:
        idx, dist, st = ctx.knn(data, query, int(k), _layout_of(data), algo, n_cand)
    if stats_out is not None:
        stats_out.update(st)
    _warn_slow_path(st, idx.shape[0])
    out = {"nn.idx": idx, "nn.dists": dist}
    if cache_index:
        out["idx"] = index
    return out


def NNHelper(data, query=None, k: int = 20, method: str = "rann", cache_index: bool = False,
             ctx: _lib.Context | None = None, knn_algo: int = _lib.KNN_AUTO, n_cand: int = 0,
             stats_out: dict | None = None, index=None, **kwargs) -> Neighbor:
    """R/clustering.R:1658-1690.  Only the exact ``"rann"`` arm is accelerated;
    ``searchtype``/``eps`` other than the exact defaults are refused.  ``cache_index`` /
    ``index`` keep / reuse the reference embedding on the device (``Neighbor.alg_idx``)."""
    data = _as_embedding(data, "data")
    query_e = data if query is None else _as_embedding(query, "query")
    if method == "rann":
        if kwargs.get("eps", 0) not in (0, 0.0) or kwargs.get("searchtype", "standard") != "standard":
            raise NotImplementedError("the B200 path implements exact search only (eps = 0, "
                                      "searchtype = 'standard')")
        if index is not None and not isinstance(index, _lib.Index):
            raise TypeError("index must be the alg_idx of a Neighbor computed with cache_index=True")
        if index is not None and (index.rows != data.values.shape[0] or index.dims != data.values.shape[1]):
            raise ValueError("the cached index does not match the dimensions of data")
        ctx = ctx or (index._ctx if index is not None else _lib.default_context())
        results = _nn2(data.values, None if query is None or query_e is data else query_e.values, k,
                       ctx, knn_algo, n_cand, stats_out, cache_index=cache_index is True, index=index)
    elif method in ("annoy", "hnsw"):
        raise NotImplementedError(f"nn.method = '{method}' is an approximate back end of the "
                                  "reference and outside the B200 path; use 'rann'")
    else:
        raise ValueError("Invalid method. Please choose one of 'rann', 'annoy'")
    n_ob = Neighbor(nn_idx=results["nn.idx"], nn_dist=results["nn.dists"],
                    alg_info=results.get("alg.info") or {}, cell_names=query_e.rownames)
    if cache_index is True and results.get("idx") is not None:
        n_ob.alg_idx = results["idx"]
    return n_ob


def ComputeSNN(nn_ranked, prune: float, ctx: _lib.Context | None = None,
               stats_out: dict | None = None) -> Graph:
    """``ComputeSNN(nn_ranked, prune)`` -- src/snn.cpp:14-38.  ``nn_ranked`` is an
    integer (or whole-number double, as Rcpp coerces) matrix of 1-based rows."""
    nn = np.asarray(nn_ranked)
    if nn.ndim != 2:
        raise ValueError("nn_ranked must be a matrix")
    if nn.dtype != np.int32:
        nn = nn.astype(np.int32)  # Rcpp's (int) cast of the double entries
    if not (nn.flags.c_contiguous or nn.flags.f_contiguous):
        nn = np.ascontiguousarray(nn)
    ctx = ctx or _lib.default_context()
    p, i, x, st = ctx.compute_snn(nn, float(prune), _layout_of(nn))
    if stats_out is not None:
        stats_out.update(st)
    n = nn.shape[0]
    return Graph(p, i, x, (n, n))


def WriteEdgeFile(snn: Graph, filename: str, display_progress: bool = False) -> None:
    """``WriteEdgeFile(snn, filename, display_progress)`` -- src/snn.cpp:40-59: the strict lower triangle
    of the graph, one ``col<TAB>row<TAB>value`` line per entry, 15 significant digits."""
    if display_progress:
        _message("Writing SNN as edge file")
    _lib.write_edge_file(snn.p, snn.i, snn.x, str(filename))


def DirectSNNToFile(nn_ranked, prune: float, display_progress: bool, filename: str,
                    ctx: _lib.Context | None = None) -> Graph:
    """``DirectSNNToFile`` -- src/snn.cpp:63-69: ComputeSNN, then the edge file, without a trip through R."""
    snn = ComputeSNN(nn_ranked, prune, ctx=ctx)
    WriteEdgeFile(snn, filename, display_progress)
    return snn


def fast_dist(x, y, n, ctx: _lib.Context | None = None) -> list:
    """``fast_dist(x, y, n)`` -- src/fast_NN_dist.cpp:34-69: for every row r of ``x`` the Euclidean distances
    to the rows of ``y`` listed (1-based) in ``n[r]``.  Returns a list of vectors shaped like ``n``;
    an empty list when ``len(n) != nrow(x)``, like the reference."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.shape[0] != len(n):
        return []
    if not (x.flags.c_contiguous or x.flags.f_contiguous):
        x = np.ascontiguousarray(x)
    if _layout_of(y) != _layout_of(x) or not (y.flags.c_contiguous or y.flags.f_contiguous):
        y = np.asarray(y, order="C" if _layout_of(x) == _lib.ROW_MAJOR else "F")
    ptr = np.zeros(len(n) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(v) for v in n])
    nbr = (np.concatenate([np.asarray(v).astype(np.int32) for v in n]) if ptr[-1] else np.zeros(0, np.int32))
    ctx = ctx or _lib.default_context()
    out = ctx.fast_dist(x, y, ptr, nbr, _layout_of(x))
    return [out[ptr[r]:ptr[r + 1]] for r in range(len(n))]


def SNN_SmallestNonzero_Dist(snn: Graph, mat, n: int, nearest_dist, ctx: _lib.Context | None = None) -> np.ndarray:
    """``SNN_SmallestNonzero_Dist(snn, mat, n, nearest_dist)`` -- src/snn.cpp:82-126 (the WNN kernel-bandwidth
    helper): per cell the mean distance to its ``n`` weakest SNN neighbours, minus ``nearest_dist``."""
    mat = np.asarray(mat, dtype=np.float64)
    if not (mat.flags.c_contiguous or mat.flags.f_contiguous):
        mat = np.ascontiguousarray(mat)
    ctx = ctx or _lib.default_context()
    return ctx.snn_smallest_nonzero_dist(snn.p, snn.i, snn.x, mat, _layout_of(mat), int(n), nearest_dist)


def ComputeSNNwidth(snn_graph: Graph, embeddings, k_nn: int, l2_norm: bool = True, nearest_dist=None,
                    ctx: _lib.Context | None = None) -> np.ndarray:
    """``ComputeSNNwidth`` (R/clustering.R:1027-1049), the R-level caller of ``SNN_SmallestNonzero_Dist``
    (WNN's kernel bandwidth): optional row L2 normalisation, ``nearest.dist`` defaults to zeros and must
    have one entry per column of the graph."""
    emb = np.asarray(embeddings, dtype=np.float64)
    if l2_norm:
        emb = L2Norm(emb)
    ncol = snn_graph.Dim[1]
    nearest = np.zeros(ncol, dtype=np.float64) if nearest_dist is None else np.asarray(nearest_dist, dtype=np.float64)
    if nearest.size != ncol:
        raise ValueError("Please provide a vector for nearest.dist that has as many elements as"
                         f" there are columns in the snn.graph ({ncol}).")
    return SNN_SmallestNonzero_Dist(snn_graph, emb, k_nn, nearest, ctx=ctx)


def NNdist(nn_idx, embeddings, metric: str = "euclidean", query_embeddings=None, nearest_dist=None,
           ctx: _lib.Context | None = None) -> list:
    """``NNdist`` (R/clustering.R:1620-1644), the R-level caller of ``fast_dist``: distances from every
    (query) cell to the cells listed for it (a matrix of 1-based indices, one row per cell, or a list of
    index vectors); with ``nearest.dist`` the distance to the nearest neighbour is subtracted and the
    result floored at 0."""
    emb = np.asarray(embeddings, dtype=np.float64)
    if not isinstance(nn_idx, (list, tuple)):
        nn_idx = list(np.asarray(nn_idx))
    q = emb if query_embeddings is None else np.asarray(query_embeddings, dtype=np.float64)
    d = fast_dist(q, emb, nn_idx, ctx=ctx)
    if nearest_dist is not None:
        nearest = np.asarray(nearest_dist, dtype=np.float64)
        d = [np.maximum(d[r] - nearest[r], 0.0) for r in range(q.shape[0])]
    return d


def _nn_graph(nn_ranked: np.ndarray, n: int, ctx: _lib.Context) -> Graph:
    nn = np.asarray(nn_ranked)
    if nn.dtype != np.int32:
        nn = nn.astype(np.int32)
    if not (nn.flags.c_contiguous or nn.flags.f_contiguous):
        nn = np.ascontiguousarray(nn)
    p, i, x, _ = ctx.nn_graph(nn, n, _layout_of(nn))
    return Graph(p, i, x, (n, n))


# ---------------------------------------------------------------------------
# FindNeighbors.default
# ---------------------------------------------------------------------------
def FindNeighbors(object, query=None, distance_matrix: bool = False, k_param: int = 20,
                  return_neighbor: bool = False, compute_SNN: bool | None = None,
                  prune_SNN: float = 1 / 15, nn_method: str = "rann", n_trees: int = 50,
                  annoy_metric: str = "euclidean", nn_eps: float = 0, verbose: bool = True,
                  l2_norm: bool = False, cache_index: bool = False, index=None,
                  ctx: _lib.Context | None = None, knn_algo: int = _lib.KNN_AUTO, n_cand: int = 0,
                  stats_out: dict | None = None):
    """``FindNeighbors.default`` (R/clustering.R:582-685) on the B200 path.

    Returns ``{"nn": Graph, "snn": Graph}``, ``{"nn": Graph}`` when
    ``compute_SNN`` is false, or a :class:`Neighbor` when ``return_neighbor``.
    ``nn_method`` defaults to the exact ``"rann"`` arm, the one this package
    replaces (the reference's default, annoy, is approximate and out of scope).
    """
    if compute_SNN is None:
        compute_SNN = not return_neighbor
    object = _as_embedding(object)
    if object.values.ndim != 2:
        warnings.warn("Object should have two dimensions, attempting to coerce to matrix")
        object = Embedding(np.atleast_2d(object.values).reshape(len(object.values), -1),
                           object.rownames)
    if object.rownames is None:
        raise ValueError("Please provide rownames (cell names) with the input object")
    n_cells = object.shape[0]
    if n_cells < k_param:
        warnings.warn("k.param set larger than number of cells. Setting k.param to number of cells - 1.")
        k_param = n_cells - 1
    if query is not None:
        query = _as_embedding(query, "query")
    if l2_norm:
        object = L2Norm(object)
        if query is not None:
            query = L2Norm(query)
    ctx = ctx or _lib.default_context()
    if not distance_matrix:
        _message("Computing nearest neighbors" if return_neighbor
                 else "Computing nearest neighbor graph", verbose)
        fused = (query is None and not return_neighbor and nn_method == "rann" and nn_eps == 0 and index is None)
        if fused:
            # kNN -> nn Graph -> SNN with the index resident on the device
            _check_finite(object.values, "data")
            if compute_SNN:
                _message("Computing SNN", verbose)
            r = ctx.find_neighbors(object.values, int(k_param), float(prune_SNN), bool(compute_SNN),
                                   _layout_of(object.values), knn_algo, n_cand, want_knn=False)
            if stats_out is not None:
                stats_out.update(r["stats"])
            _warn_slow_path(r["stats"], n_cells)
            names = object.rownames
            graphs = {"nn": Graph(*r["nn"], (n_cells, n_cells), (names, names))}
            if compute_SNN:
                graphs["snn"] = Graph(*r["snn"], (n_cells, n_cells), (names, names))
            return graphs
        nn_ranked = NNHelper(data=object, query=query, k=k_param, method=nn_method,
                             n_trees=n_trees, searchtype="standard", eps=nn_eps,
                             metric=annoy_metric, cache_index=cache_index, index=index, ctx=ctx,
                             knn_algo=knn_algo, n_cand=n_cand, stats_out=stats_out)
        if return_neighbor:
            if compute_SNN:
                warnings.warn("The SNN graph is not computed if return.neighbor is TRUE.")
            return nn_ranked
        nn_ranked = Indices(nn_ranked)
    else:
        _message("Building SNN based on a provided distance matrix", verbose)
        # R/clustering.R:655-662: order() of every row, first k.param columns
        d = object.values
        knn_mat = np.argsort(d, axis=1, kind="stable")[:, :k_param].astype(np.int32) + 1
        nn_ranked = knn_mat
    # nn.ranked -> Graph (R/clustering.R:664-670); dims are nrow(object) squared
    nn_matrix = _nn_graph(nn_ranked, n_cells, ctx)
    nn_matrix.rownames = object.rownames
    nn_matrix.colnames = object.rownames
    neighbor_graphs = {"nn": nn_matrix}
    if compute_SNN:
        _message("Computing SNN", verbose)
        snn_matrix = ComputeSNN(nn_ranked=nn_ranked, prune=prune_SNN, ctx=ctx, stats_out=stats_out)
        snn_matrix.rownames = object.rownames
        snn_matrix.colnames = object.rownames
        neighbor_graphs["snn"] = as_Graph(snn_matrix)
    return neighbor_graphs


# ---------------------------------------------------------------------------
# FindNeighbors on an embedding that already lives in HBM
# ---------------------------------------------------------------------------
def FindNeighborsResident(embedding, rownames, dims=None, k_param: int = 20, return_neighbor: bool = False,
                          compute_SNN: bool | None = None, prune_SNN: float = 1 / 15, l2_norm: bool = False,
                          verbose: bool = True, ctx: _lib.Context | None = None, stats_out: dict | None = None):
    """``FindNeighbors`` for a device-resident embedding -- e.g. a PCA computed on the GPU: the 400 MB - 1 GB
    upload of the embedding and the host-side ``[, dims]`` / ``L2Norm`` copies disappear.

    ``embedding``: 2-D float64 CUDA tensor [cells x components], row-major (contiguous) or column-major (the
    transpose is contiguous, R's layout); ``dims``: 0-based component numbers (``dims = 1:10`` of
    ``FindNeighbors.Seurat``, R/clustering.R:843-846, is ``range(10)``), None = all; ``l2_norm`` as in
    ``FindNeighbors.default`` (R/clustering.R:619-622), R's arithmetic bit for bit.  Same return values as
    :func:`FindNeighbors`: ``{"nn": Graph, "snn": Graph}`` / ``{"nn": Graph}`` / :class:`Neighbor`."""
    import torch
    t = embedding
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.dim() == 2):
        raise TypeError("FindNeighborsResident expects a 2-D float64 CUDA tensor; use FindNeighbors for host matrices")
    if rownames is None:
        raise ValueError("Please provide rownames (cell names) with the input object")
    if compute_SNN is None:
        compute_SNN = not return_neighbor
    n, n_cols = int(t.shape[0]), int(t.shape[1])
    if len(rownames) != n:
        raise ValueError("rownames must name every row of the embedding")
    if t.is_contiguous():
        layout = _lib.ROW_MAJOR
    elif t.t().is_contiguous():
        layout = _lib.COL_MAJOR
    else:
        t, layout = t.contiguous(), _lib.ROW_MAJOR
    dims0 = None if dims is None else np.asarray(list(dims), dtype=np.int32)
    nd = n_cols if dims0 is None else int(dims0.size)
    if n < k_param:
        warnings.warn("k.param set larger than number of cells. Setting k.param to number of cells - 1.")
        k_param = n - 1
    ctx = ctx or _lib.default_context()
    with torch.cuda.device(t.device):
        if dims0 is None and not l2_norm and layout == _lib.ROW_MAJOR:
            work = t                                        # already what the search takes
        else:
            work = torch.empty((n, nd), dtype=torch.float64, device=t.device)
            ctx.dev_embedding_prepare(t, n, n_cols, layout, dims0, l2_norm, work)
        _message("Computing nearest neighbors" if return_neighbor else "Computing nearest neighbor graph", verbose)
        idx0 = torch.empty((n, k_param), dtype=torch.int32, device=t.device)
        dist = torch.empty((n, k_param), dtype=torch.float64, device=t.device)
        st = ctx.dev_knn(work, n, None, n, nd, int(k_param), idx0, dist)   # refuses NA / NaN / Inf itself
        _warn_slow_path(st, n)
        if return_neighbor:
            if compute_SNN:
                warnings.warn("The SNN graph is not computed if return.neighbor is TRUE.")
            if stats_out is not None:
                stats_out.update(st)
            return Neighbor(nn_idx=(idx0 + 1).cpu().numpy(), nn_dist=dist.cpu().numpy(), cell_names=list(rownames))
        if compute_SNN:
            _message("Computing SNN", verbose)
        g = ctx.dev_graphs(idx0, n, int(k_param), float(prune_SNN), bool(compute_SNN), 0, n)
        if stats_out is not None:
            stats_out.update(st)
            stats_out.update(g)
        out = {}
        for name, which in (("nn", 0), ("snn", 1)) if compute_SNN else (("nn", 0),):
            r = ctx.dev_result(which)
            P = torch.empty(r.n_rows + 1, dtype=torch.int64, device=t.device)
            I = torch.empty(r.nnz, dtype=torch.int32, device=t.device)
            X = torch.empty(r.nnz, dtype=torch.float64, device=t.device)
            ctx.dev_copy_result(which, P, I, X)
            out[name] = Graph(P.to(torch.int32).cpu().numpy(), I.cpu().numpy(), X.cpu().numpy(), (n, n),
                              (list(rownames), list(rownames)))
    return out


# ---------------------------------------------------------------------------
# FindNeighbors.Seurat on a minimal object
# ---------------------------------------------------------------------------
class SeuratLite:
    """Just enough of a Seurat object for ``FindNeighbors.Seurat``: named
    dimensional reductions (cell embeddings), a default assay name, and the
    ``graphs`` / ``neighbors`` / ``commands`` slots the method writes."""

    def __init__(self, reductions: dict, default_assay: str = "RNA"):
        self.reductions = {k: _as_embedding(v) for k, v in reductions.items()}
        self.default_assay = default_assay
        self.graphs: dict = {}
        self.neighbors: dict = {}
        self.commands: list = []


def FindNeighborsSeurat(object: SeuratLite, reduction: str = "pca", dims=range(10), assay=None,
                        k_param: int = 20, return_neighbor: bool = False,
                        compute_SNN: bool | None = None, prune_SNN: float = 1 / 15,
                        nn_method: str = "rann", nn_eps: float = 0, verbose: bool = True,
                        graph_name=None, l2_norm: bool = False, cache_index: bool = False, **kw):
    """``FindNeighbors.Seurat`` (R/clustering.R:790-898), reduction branch.
    ``dims`` are 0-based Python positions (R's ``1:10`` is ``range(10)``)."""
    if compute_SNN is None:
        compute_SNN = not return_neighbor
    if reduction not in object.reductions:
        raise KeyError(f"reduction '{reduction}' not found")
    assay = assay or object.default_assay
    emb = object.reductions[reduction]
    dims = list(dims)
    if max(dims) >= emb.shape[1]:
        raise ValueError("More dimensions specified in dims than have been computed")
    data_use = emb.take_columns(dims)
    res = FindNeighbors(data_use, k_param=k_param, return_neighbor=return_neighbor,
                        compute_SNN=compute_SNN, prune_SNN=prune_SNN, nn_method=nn_method,
                        nn_eps=nn_eps, verbose=verbose, l2_norm=l2_norm, cache_index=cache_index,
                        **kw)
    if isinstance(res, Neighbor):
        res = {"nn": res}
    if graph_name is None:
        sep = "." if return_neighbor else "_"
        names = [f"{assay}{sep}{n}" for n in res.keys()]
    else:
        names = [graph_name] if isinstance(graph_name, str) else list(graph_name)
    if len(names) == 1:
        _message("Only one graph name supplied, storing nearest-neighbor graph only", True)
    for name, g in zip(names, res.values()):
        if isinstance(g, Graph):
            g.assay_used = assay
            object.graphs[name] = g
        else:
            object.neighbors[name] = g
    object.commands.append({"name": "FindNeighbors", "reduction": reduction, "dims": dims,
                            "k.param": k_param, "prune.SNN": prune_SNN, "nn.method": nn_method})
    return object
