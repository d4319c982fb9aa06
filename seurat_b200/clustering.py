"""Host mirror of the clustering step that follows FindNeighbors.

Mirrors, argument for argument, the reference's R functions
  * ``RunModularityClustering``   R/clustering.R:1881-1906  (-> RunModularityClusteringCpp,
                                  src/RModularityOptimizer.cpp:21-173)
  * ``FindClusters.default``      R/clustering.R:307-432
  * ``GroupSingletons``           R/clustering.R:1373-1415
over the C ABI (``b200nn_modularity_cluster*``).  All compute is in libb200nn.so; there is no
CPU fallback.  Algorithms 1 (Louvain), 2 (Louvain with multilevel refinement) and 3 (smart local
moving) run on the device and give the reference's clusters for the same ``random.seed``; 4
(Leiden, an R package in the reference) is refused with an error.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .neighbors import Graph, _message

__all__ = ["RunModularityClustering", "FindClusters", "GroupSingletons"]


def _slots(SNN):
    if isinstance(SNN, Graph):
        return SNN.p, SNN.i, SNN.x, SNN.Dim[1], SNN.colnames
    if hasattr(SNN, "indptr"):  # scipy CSC / CSR (the SNN is symmetric; R hands over a dgCMatrix)
        m = SNN.tocsc()
        m.sort_indices()
        return m.indptr, m.indices, m.data, m.shape[1], None
    p, i, x = SNN
    return p, i, x, len(p) - 1, None


def RunModularityClustering(SNN, modularity: int = 1, resolution: float = 0.8, algorithm: int = 1,
                            n_start: int = 10, n_iter: int = 10, random_seed: int = 0,
                            print_output: bool = True, temp_file_location=None, edge_file_name=None,
                            ctx: _lib.Context | None = None, stats_out: dict | None = None,
                            resident: bool = False) -> np.ndarray:
    """``RunModularityClustering`` (R/clustering.R:1881-1906): cluster ids (0-based, largest cluster
    first) of every column of the SNN graph.  ``resident=True`` clusters the SNN graph the context
    still holds in HBM from the preceding FindNeighbors call instead of uploading ``SNN`` again."""
    ctx = ctx or _lib.default_context()
    if edge_file_name:
        # edge_file <- edge.file.name %||% '' (R/clustering.R:1893): a non-empty name replaces the matrix
        # (RModularityOptimizer.cpp:55-63); the number of nodes is then the largest id in the file + 1
        first = ctx._first() if hasattr(ctx, "_first") else ctx
        ids, st = first.modularity_cluster_edge_file(edge_file_name, modularity, resolution, algorithm, n_start,
                                                     n_iter, random_seed)
        if print_output:
            _message("Reading input file...\n")
        p = i = x = None
        n = ids.size
    else:
        p, i, x, n, _ = _slots(SNN)
    if edge_file_name:
        pass
    elif resident:
        ids, st = ctx.modularity_cluster_resident(n, modularity, resolution, algorithm, n_start, n_iter, random_seed)
    else:
        ids, st = ctx.modularity_cluster(p, i, x, modularity, resolution, algorithm, n_start, n_iter, random_seed)
    if stats_out is not None:
        stats_out.update(st)
    if print_output:  # the wrapper's Rprintf lines (RModularityOptimizer.cpp:88-96, 152-163)
        _message("Modularity Optimizer version 1.3.0 by Ludo Waltman and Nees Jan van Eck\n")
        _message(f"Number of nodes: {st['n_nodes']}\nNumber of edges: {st['n_edges']}\n")
        name = ("Louvain algorithm" if algorithm == 1 else "Louvain algorithm with multilevel refinement"
                if algorithm == 2 else "smart local moving algorithm")
        _message(f"Running {name}...")
        if n_start == 1:
            _message(f"Modularity: {st['modularity']:.4f}")
        else:
            _message(f"Maximum modularity in {n_start} random starts: {st['modularity']:.4f}")
        _message(f"Number of communities: {st['n_clusters']}")
        _message(f"Elapsed time: {int(st['ms_total'] / 1000.0)} seconds")
    return ids


def GroupSingletons(ids, SNN, group_singletons: bool = True, verbose: bool = True):
    """``GroupSingletons`` (R/clustering.R:1373-1415): clusters of one cell are merged into the
    cluster they are most connected to (mean SNN weight between the cell and the cluster).
    ``ids`` is a sequence of labels; returns a list of labels (strings when ungrouped singletons are
    marked "singleton", like the reference)."""
    ids = list(ids)
    labels, counts = np.unique(np.asarray(ids, dtype=object).astype(str), return_counts=True)
    single = set(labels[counts == 1])
    order = []  # unique(ids) in order of appearance
    seen = set()
    for v in ids:
        s = str(v)
        if s not in seen:
            seen.add(s)
            order.append(s)
    singletons = [s for s in order if s in single]
    if not group_singletons:
        return ["singleton" if str(v) in single else v for v in ids]
    if not singletons:
        return ids
    p, i, x, n, _ = _slots(SNN)
    import scipy.sparse as sp
    S = sp.csc_matrix((x, i, p), shape=(n, n))
    cluster_names = [s for s in order if s not in single]
    if not cluster_names:
        return ids
    new_ids = list(ids)
    str_ids = np.array([str(v) for v in ids], dtype=object)
    for s in singletons:
        cell = int(np.nonzero(str_ids == s)[0][0])
        row = S[cell, :].toarray().ravel()
        # `ids` is updated inside the loop (:1405), so a singleton merged earlier already counts as a
        # member of its new cluster here
        conn = []
        for c in cluster_names:
            cols = np.nonzero(str_ids == c)[0]
            conn.append(row[cols].sum() / (1 * len(cols)) if len(cols) else np.nan)
        conn = np.array(conn)
        m = np.nanmax(conn)
        best = [c for c, v in zip(cluster_names, conn) if v == m]
        if len(best) > 1 and verbose:
            # the reference draws among tied clusters with R's sample() after set.seed(1); R's RNG is
            # not reproduced here, the first tied cluster is taken
            _message(f"GroupSingletons: {len(best)} clusters tie for singleton {s}; taking the first")
        target = best[0]
        str_ids[cell] = target
        new_ids[cell] = target if isinstance(ids[cell], str) else type(ids[cell])(target)
    if verbose:
        _message(f"{len(singletons)} singletons identified. {len(set(map(str, new_ids)))} final clusters.")
    return new_ids


def FindClusters(object, modularity_fxn: int = 1, initial_membership=None, node_sizes=None, resolution=0.8,
                 algorithm=1, n_start: int = 10, n_iter: int = 10, random_seed: int = 0,
                 group_singletons: bool = True, temp_file_location=None, edge_file_name=None,
                 verbose: bool = True, ctx: _lib.Context | None = None, resident: bool = False) -> dict:
    """``FindClusters.default`` (R/clustering.R:307-432) on an SNN Graph: one column of cluster ids
    per resolution, keyed ``res.<r>`` like the reference's data frame."""
    if object is None:
        raise ValueError("Please provide an SNN graph")
    if isinstance(algorithm, str):
        algorithm = {"louvain": 1, "leiden": 4}.get(algorithm.lower(), algorithm)
    if algorithm not in (1, 2, 3, 4):
        raise ValueError("algorithm not recognised, please specify as an integer or string")
    res_list = list(resolution) if isinstance(resolution, (list, tuple, np.ndarray)) else [resolution]
    out = {}
    for r in res_list:
        ids = RunModularityClustering(SNN=object, modularity=modularity_fxn, resolution=r, algorithm=algorithm,
                                      n_start=n_start, n_iter=n_iter, random_seed=random_seed, print_output=verbose,
                                      temp_file_location=temp_file_location, edge_file_name=edge_file_name, ctx=ctx,
                                      resident=resident)
        ids = GroupSingletons(ids.tolist(), object, group_singletons=group_singletons, verbose=verbose)
        out[f"res.{r}"] = np.asarray(ids)
    return out
