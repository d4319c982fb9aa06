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


# ---------------------------------------------------------------------------
# R's random draw, for the one place the reference's result depends on it: GroupSingletons calls
# set.seed(1) and then sample(tied clusters, 1) (R/clustering.R:1395-1405).  Restated from R's
# sources (src/main/RNG.c: set.seed's scrambling and the Mersenne-Twister state; R_unif_index with
# sample.kind = "Rejection", R >= 3.6); pinned in tests/test_modopt_cpu.py by R's well-known outputs
# (set.seed(1); runif(3) -> 0.2655087 0.3721239 0.5728534; sample(1:10) -> 9 4 7 1 2 5 3 10 6 8).
# ---------------------------------------------------------------------------
def _r_unif_rand(seed: int):
    """generator of unif_rand() values after set.seed(seed) (default RNG kind "Mersenne-Twister")"""
    m32 = 0xffffffff
    s = int(seed) & m32
    for _ in range(50):                      # RNG_Init: initial scrambling
        s = (69069 * s + 1) & m32
    st = []
    for _ in range(625):
        s = (69069 * s + 1) & m32
        st.append(s)
    mt = st[1:]                              # dummy[0] = mti = 624: the first draw regenerates the state
    n_, m_ = 624, 397
    mti = n_
    while True:
        if mti >= n_:
            for kk in range(n_):
                y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % n_] & 0x7fffffff)
                mt[kk] = mt[(kk + m_) % n_] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
            mti = 0
        y = mt[mti]
        mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        v = y * 2.3283064365386963e-10       # [0, 1), then fixup() keeps it inside (0, 1)
        if v <= 0.0:
            v = 0.5 * 2.328306437080797e-10
        elif 1.0 - v <= 0.0:
            v = 1.0 - 0.5 * 2.328306437080797e-10
        yield v


def _r_unif_index(gen, dn: int) -> int:
    """R_unif_index(dn), rejection sampling from the integers below the next power of two"""
    if dn <= 0:
        return 0
    bits = (dn - 1).bit_length()             # ceil(log2(dn))
    while True:
        v, n = 0, 0
        while n <= bits:                     # rbits(): 16 random bits per unif_rand()
            v = 65536 * v + int(next(gen) * 65536)
            n += 16
        v &= (1 << bits) - 1
        if v < dn:
            return v


def _r_mean(x) -> float:
    """R's mean() of a double vector (src/main/summary.c, real_mean): the sum in long double in storage
    order, divided by n, then one refinement pass s += sum(x - s) / n; rounded to double at the end.
    GroupSingletons compares such means for equality, so the arithmetic matters."""
    v = np.asarray(x, dtype=np.float64).astype(np.longdouble)
    n = np.longdouble(v.size)
    s = np.cumsum(v)[-1] / n                 # cumsum adds term by term (np.sum is pairwise)
    s = s + np.cumsum(v - s)[-1] / n
    return float(s)


def _r_sample(n: int, size: int | None = None, seed: int = 1) -> list:
    """sample.int(n, size) without replacement right after set.seed(seed): 0-based positions"""
    gen = _r_unif_rand(seed)
    x = list(range(n))
    out = []
    for _ in range(n if size is None else size):
        j = _r_unif_index(gen, n)
        out.append(x[j])
        n -= 1
        x[j] = x[n]
    return out


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
            conn.append(_r_mean(row[cols]) if len(cols) else np.nan)
        conn = np.array(conn)
        m = np.nanmax(conn)
        best = [c for c, v in zip(cluster_names, conn) if v == m]
        # closest_cluster <- sample(names(connectivity[mi]), 1) right after set.seed(1) (:1395, :1404): R's
        # draw, reproduced (the first of up to 8 tied clusters, the 9th of 9 - 16, ...)
        target = best[_r_sample(len(best), 1, seed=1)[0]]
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
