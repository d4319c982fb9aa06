"""GPU suite (-m gpu): the CUDA path, called through the C ABI (ctypes), against the
oracle and the committed golden fixtures.

Bars (north_star): kNN indices identical (positions whose oracle distances tie within
1e-6 relative are interchangeable), distances within 1e-5 relative -- in fact the FP64
re-rank reproduces the oracle's distances bit for bit, which is asserted where no tie
handling is involved; nn Graph and SNN (p, i, x) bit-exact given identical kNN.
"""
import os

import numpy as np
import pytest

import oracle
import seurat_b200 as sb
from conftest import assert_csr_equal, assert_knn_matches, unhex
from seurat_b200 import _lib
from seurat_b200.synthetic import cell_names, gaussian_mixture

pytestmark = pytest.mark.gpu

ALGOS = [_lib.KNN_EXACT, _lib.KNN_AUTO]


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def emb(X, order="C"):
    return sb.Embedding(np.asarray(X, order=order), cell_names(len(X)))


# ---------------------------------------------------------------- golden vectors
def test_survey_known_answers(ctx, fixtures):
    s = fixtures["survey_n6"]
    nn = np.array(s["nn_ranked"], dtype=np.int32)
    for key, prune in (("prune_0", 0.0), ("prune_1/15", 1 / 15), ("prune_0.21", 0.21)):
        want = (np.array(s[key]["p"]), np.array(s[key]["i"]), np.array(s[key]["x"], dtype=np.float64))
        for order in ("C", "F"):
            g = sb.ComputeSNN(np.asarray(nn, order=order), prune, ctx=ctx)
            assert_csr_equal((g.p, g.i, g.x), want, f"{key} {order}")
            assert g.Dim == (6, 6)


def test_snn_edge_cases(ctx, fixtures):
    for name, c in fixtures["snn_cases"].items():
        nn = np.array(c["nn_ranked"], dtype=np.int32)
        prune = float.fromhex(c["prune"])
        want = (np.array(c["snn"]["p"]), np.array(c["snn"]["i"]), unhex(c["snn"]["x"]))
        g = sb.ComputeSNN(np.asfortranarray(nn), prune, ctx=ctx)
        assert_csr_equal((g.p, g.i, g.x), want, name)
        # doubles holding whole numbers are accepted like Rcpp's coercion
        g2 = sb.ComputeSNN(nn.astype(np.float64), prune, ctx=ctx)
        assert_csr_equal((g2.p, g2.i, g2.x), want, name + " (double input)")
        wantg = (np.array(c["nn"]["p"]), np.array(c["nn"]["i"]), unhex(c["nn"]["x"]))
        p, i, x, _ = ctx.nn_graph(nn, nn.shape[0], _lib.ROW_MAJOR)
        assert_csr_equal((p, i, x), wantg, name + " nn graph")


def test_snn_index_out_of_range(ctx):
    nn = np.array([[1, 2], [2, 4], [3, 1]], dtype=np.int32)  # 4 > n = 3
    with pytest.raises(sb.B200Error) as ei:
        sb.ComputeSNN(nn, 0.0, ctx=ctx)
    assert ei.value.code == _lib.ERR_INDEX_RANGE
    nn[1, 1] = 0
    with pytest.raises(sb.B200Error):
        sb.ComputeSNN(nn, 0.0, ctx=ctx)


@pytest.mark.parametrize("algo", ALGOS)
def test_pbmc_config_a(ctx, pbmc, algo):
    """config A: pbmc-sized 80 x 10, k = 20 (tests/golden/make_golden.py)"""
    for order in ("C", "F"):
        X = emb(pbmc["X"], order)
        nb = sb.FindNeighbors(X, k_param=20, return_neighbor=True, verbose=False, ctx=ctx, knn_algo=algo)
        assert nb.cell_names == X.rownames
        assert (nb.nn_idx == pbmc["nn_idx"]).all()
        assert (nb.nn_dist == pbmc["nn_dist"]).all()  # FP64 re-rank: bit-exact
        g = sb.FindNeighbors(X, k_param=20, prune_SNN=pbmc["prune"], verbose=False, ctx=ctx, knn_algo=algo)
        assert_csr_equal((g["nn"].p, g["nn"].i, g["nn"].x), pbmc["nn"], "nn")
        assert_csr_equal((g["snn"].p, g["snn"].i, g["snn"].x), pbmc["snn"], "snn")
        assert g["snn"].rownames == X.rownames and g["snn"].colnames == X.rownames
        # unfused route (NNHelper -> nn graph -> ComputeSNN) gives the same graphs
        snn = sb.ComputeSNN(nb.nn_idx, pbmc["prune"], ctx=ctx)
        assert_csr_equal((snn.p, snn.i, snn.x), pbmc["snn"], "snn unfused")


@pytest.mark.parametrize("algo", ALGOS)
def test_pbmc_query_path(ctx, fixtures, pbmc, algo):
    q = fixtures["pbmc_query"]
    X = pbmc["X"]
    ref = emb(X[q["ref_rows"][0]:q["ref_rows"][1]])
    qry = sb.Embedding(np.ascontiguousarray(X[q["query_rows"][0]:q["query_rows"][1]]),
                       pbmc["cells"][q["query_rows"][0]:q["query_rows"][1]])
    nb = sb.FindNeighbors(ref, query=qry, k_param=q["k"], return_neighbor=True, verbose=False, ctx=ctx,
                          knn_algo=algo)
    assert nb.cell_names == qry.rownames
    assert (nb.nn_idx == np.array(q["nn_idx"])).all()
    from conftest import unhex2
    assert (nb.nn_dist == unhex2(q["nn_dist"])).all()


# ---------------------------------------------------------------- host semantics on the GPU path
def test_k_clipped_with_warning(ctx):
    X = emb(np.random.default_rng(0).normal(size=(12, 4)))
    with pytest.warns(UserWarning, match="k.param set larger than number of cells"):
        nb = sb.FindNeighbors(X, k_param=20, return_neighbor=True, verbose=False, ctx=ctx)
    assert nb.nn_idx.shape == (12, 11)
    wi, wd = oracle.knn(X.values, k=11)
    assert (nb.nn_idx == wi).all() and (nb.nn_dist == wd).all()


def test_k_larger_than_reference_errors(ctx):
    X = np.random.default_rng(0).normal(size=(5, 3))
    with pytest.raises(ValueError, match="Cannot find more nearest neighbours"):
        sb.NNHelper(emb(X), k=6, method="rann", ctx=ctx)
    with pytest.raises(sb.B200Error) as ei:
        ctx.knn(X, None, 6, _lib.ROW_MAJOR)
    assert ei.value.code == _lib.ERR_K_TOO_LARGE


def test_return_neighbor_with_compute_snn_warns(ctx):
    X = emb(np.random.default_rng(1).normal(size=(40, 5)))
    with pytest.warns(UserWarning, match="SNN graph is not computed"):
        nb = sb.FindNeighbors(X, k_param=5, return_neighbor=True, compute_SNN=True, verbose=False, ctx=ctx)
    assert isinstance(nb, sb.Neighbor)


def test_compute_snn_false_returns_nn_only(ctx):
    X = emb(np.random.default_rng(1).normal(size=(40, 5)))
    g = sb.FindNeighbors(X, k_param=5, compute_SNN=False, verbose=False, ctx=ctx)
    assert list(g.keys()) == ["nn"]
    wi, _ = oracle.knn(X.values, k=5)
    assert_csr_equal((g["nn"].p, g["nn"].i, g["nn"].x), oracle.nn_graph(wi), "nn only")


def test_l2_norm_option(ctx):
    X = np.random.default_rng(3).normal(size=(60, 6)) * 5
    g = sb.FindNeighbors(emb(X), k_param=8, l2_norm=True, verbose=False, ctx=ctx)
    w = oracle.find_neighbors(oracle.l2norm(X), k=8)
    assert_csr_equal((g["snn"].p, g["snn"].i, g["snn"].x), w["snn"], "l2 snn")


def test_query_ne_object_graph_quirk(ctx):
    """R/clustering.R:665-667 builds an nrow(object)^2 matrix from the query's rows"""
    rng = np.random.default_rng(4)
    obj, qry = rng.normal(size=(50, 4)), rng.normal(size=(50, 4))
    g = sb.FindNeighbors(emb(obj), query=sb.Embedding(qry, cell_names(50)), k_param=6, verbose=False, ctx=ctx)
    wi, _ = oracle.knn(obj, qry, 6)
    assert_csr_equal((g["nn"].p, g["nn"].i, g["nn"].x), oracle.nn_graph(wi, 50), "quirk nn")
    assert_csr_equal((g["snn"].p, g["snn"].i, g["snn"].x), oracle.compute_snn(wi, 1 / 15), "quirk snn")


def test_distance_matrix_input(ctx):
    rng = np.random.default_rng(8)
    P = rng.normal(size=(30, 3))
    D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
    g = sb.FindNeighbors(sb.Embedding(D, cell_names(30)), distance_matrix=True, k_param=5, verbose=False, ctx=ctx)
    knn = np.argsort(D, axis=1, kind="stable")[:, :5] + 1
    assert_csr_equal((g["snn"].p, g["snn"].i, g["snn"].x), oracle.compute_snn(knn, 1 / 15), "distmat snn")


def test_seurat_method(ctx, pbmc):
    X = np.hstack([pbmc["X"], np.random.default_rng(0).normal(size=(80, 5))])
    so = sb.SeuratLite({"pca": sb.Embedding(X, pbmc["cells"])}, default_assay="RNA")
    so = sb.FindNeighborsSeurat(so, dims=range(10), k_param=20, verbose=False, ctx=ctx)
    assert set(so.graphs) == {"RNA_nn", "RNA_snn"}
    assert so.graphs["RNA_snn"].assay_used == "RNA"
    assert_csr_equal((so.graphs["RNA_snn"].p, so.graphs["RNA_snn"].i, so.graphs["RNA_snn"].x), pbmc["snn"], "seurat")
    so = sb.FindNeighborsSeurat(so, dims=range(10), k_param=20, return_neighbor=True, verbose=False, ctx=ctx)
    assert "RNA.nn" in so.neighbors and (so.neighbors["RNA.nn"].nn_idx == pbmc["nn_idx"]).all()
    with pytest.raises(ValueError, match="More dimensions specified"):
        sb.FindNeighborsSeurat(so, dims=range(40), verbose=False, ctx=ctx)


# ---------------------------------------------------------------- seeded parity vs the oracle
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n,d,k", [(1000, 10, 20), (5000, 50, 20), (3000, 30, 30), (257, 3, 5),
                                   (2048, 64, 10), (700, 100, 15), (4000, 50, 1), (129, 50, 128)])
def test_knn_parity_random(ctx, algo, n, d, k):
    X, _ = gaussian_mixture(n, d, seed=n + d + k, n_centres=8)
    wi, wd = oracle.knn(X, k=k)
    for order in ("C", "F"):
        idx, dist, st = ctx.knn(np.asarray(X, order=order), None, k,
                                _lib.ROW_MAJOR if order == "C" else _lib.COL_MAJOR, algo)
        assert_knn_matches(idx, dist, wi, wd, f"n={n} d={d} k={k} {order}")
        assert (dist == wd).all(), "FP64 re-rank must reproduce the oracle's distances bit for bit"
        assert (idx == wi).all()


@pytest.mark.parametrize("algo", ALGOS)
def test_knn_query_path_random(ctx, algo):
    ref, c = gaussian_mixture(6000, 50, seed=3, n_centres=8)
    qry, _ = gaussian_mixture(1500, 50, seed=4, centres=c)
    wi, wd = oracle.knn(ref, qry, 20)
    idx, dist, _ = ctx.knn(ref, qry, 20, _lib.ROW_MAJOR, algo)
    assert (idx == wi).all() and (dist == wd).all()


@pytest.mark.parametrize("algo", ALGOS)
def test_knn_ties_and_duplicates(ctx, algo):
    """exact duplicates and lattice points: ties must come out ordered by lower index"""
    rng = np.random.default_rng(9)
    base = rng.integers(0, 3, size=(300, 4)).astype(np.float64)   # many exact ties
    X = np.vstack([base, base[:50]])                              # and exact duplicates
    wi, wd = oracle.knn(X, k=12)
    idx, dist, _ = ctx.knn(X, None, 12, _lib.ROW_MAJOR, algo)
    assert (dist == wd).all()
    assert (idx == wi).all()
    Z = np.zeros((40, 50))
    idx, dist, _ = ctx.knn(Z, None, 7, _lib.ROW_MAJOR, algo)
    assert (idx == np.tile(np.arange(1, 8), (40, 1))).all() and (dist == 0).all()


@pytest.mark.parametrize("prune", [0.0, 1 / 15, 0.25, 1.0])
def test_snn_parity_random(ctx, prune):
    X, _ = gaussian_mixture(20000, 50, seed=5)
    idx, _ = oracle.knn(X, k=20, method="kdtree")
    want = oracle.compute_snn(idx, prune)
    g = sb.ComputeSNN(idx, prune, ctx=ctx)
    assert_csr_equal((g.p, g.i, g.x), want, f"prune={prune}")


def test_snn_hub_rows_all_classes(ctx):
    """skewed lists so that every work class (warp / half-block / block / HBM-hash rows,
    warp / block / in-place column sorts) is exercised"""
    rng = np.random.default_rng(6)
    n, k = 60000, 12
    nn = rng.integers(1, n + 1, size=(n, k)).astype(np.int32)
    nn[:, 0] = np.arange(1, n + 1)
    nn[: n // 2, 1] = 7                      # hub with in-degree n/2 = 30000 (block sort in smem)
    nn[: 50000, 2] = 11                      # in-degree 50000 > 32768 (in-place HBM sort)
    nn[rng.random(n) < 0.05, 3] = 13         # ~3000 (block sort)
    nn[rng.random(n) < 0.005, 4] = 17        # ~300 (warp smem sort)
    stats = {}
    g = sb.ComputeSNN(nn, 1 / 15, ctx=ctx, stats_out=stats)
    want = oracle.compute_snn(nn, 1 / 15)
    assert_csr_equal((g.p, g.i, g.x), want, "hub rows")
    assert stats["max_indegree"] >= 50000
    p, i, x, _ = ctx.nn_graph(nn, n, _lib.ROW_MAJOR)
    assert_csr_equal((p, i, x), oracle.nn_graph(nn), "hub nn graph")


def test_snn_64bit_words(ctx):
    """k^2 with duplicates no longer fits the packed 32-bit word: the 64-bit table path"""
    rng = np.random.default_rng(10)
    n, k = 70000, 40          # 17 key bits -> 15 count bits < k*k = 1600?  no: force via duplicates + big k
    nn = rng.integers(1, n + 1, size=(n, k)).astype(np.int32)
    nn[:, 0] = np.arange(1, n + 1)
    nn[5, 1] = nn[5, 2]       # one duplicate switches the count range to k^2
    g = sb.ComputeSNN(nn, 1 / 15, ctx=ctx)
    assert_csr_equal((g.p, g.i, g.x), oracle.compute_snn(nn, 1 / 15), "dups large k")
    n, k = 3000, 200          # 12 key bits, k^2 = 40000 needs 16 bits; with n = 2^20+ it would spill to 64 bit
    nn = rng.integers(1, n + 1, size=(n, k)).astype(np.int32)
    nn[7, 1] = nn[7, 2]
    g = sb.ComputeSNN(nn, 0.0, ctx=ctx)
    assert_csr_equal((g.p, g.i, g.x), oracle.compute_snn(nn, 0.0), "k=200 prune 0")


@pytest.mark.parametrize("algo", ALGOS)
def test_find_neighbors_config_b_subsample(ctx, algo):
    """config B's generator at 30k cells (the oracle finishes in seconds): whole pipeline"""
    X, _ = gaussian_mixture(30000, 50, seed=0)
    want = oracle.find_neighbors(X, k=20, prune=1 / 15, method="kdtree")
    stats = {}
    g = sb.FindNeighbors(emb(X), k_param=20, verbose=False, ctx=ctx, knn_algo=algo, stats_out=stats)
    assert_csr_equal((g["nn"].p, g["nn"].i, g["nn"].x), want["nn"], "nn")
    assert_csr_equal((g["snn"].p, g["snn"].i, g["snn"].x), want["snn"], "snn")
    nb = sb.FindNeighbors(emb(X), k_param=20, return_neighbor=True, verbose=False, ctx=ctx, knn_algo=algo)
    assert (nb.nn_idx == want["nn_idx"]).all() and (nb.nn_dist == want["nn_dist"]).all()


# ---------------------------------------------------------------- full-size properties
def test_full_size_properties_config_b(ctx):
    """config B at full size (100k x 50, k = 20): size-independent properties + a sampled oracle check"""
    X, _ = gaussian_mixture(100_000, 50, seed=0)
    n, k = X.shape[0], 20
    r = ctx.find_neighbors(X, k, 1 / 15, True, _lib.ROW_MAJOR)
    idx, dist = r["idx"], r["dist"]
    assert (idx[:, 0] == np.arange(1, n + 1)).all() and (dist[:, 0] == 0).all()      # self first
    assert (np.diff(dist, axis=1) >= 0).all()                                          # ascending
    assert idx.min() >= 1 and idx.max() <= n
    assert (np.sort(idx, axis=1)[:, 1:] != np.sort(idx, axis=1)[:, :-1]).all()          # no repeats in a row
    rows = np.random.default_rng(0).choice(n, 300, replace=False)
    wi, wd = oracle.knn(X, X[rows], k, method="kdtree")
    assert (idx[rows] == wi).all() and (dist[rows] == wd).all()
    # SNN: symmetric, unit diagonal, values in the 2k-value Jaccard set, sorted, row sums of nn graph = k
    import scipy.sparse as sp
    p, i, x = r["snn"]
    S = sp.csc_matrix((x, i, p), shape=(n, n))
    assert (S != S.T).nnz == 0
    assert (S.diagonal() == 1.0).all()
    allowed = np.array([s / (k + (k - s)) for s in range(1, k + 1)])
    assert np.isin(x, allowed[allowed >= 1 / 15]).all()
    assert all((np.diff(i[p[c]:p[c + 1]]) > 0).all() for c in rows)
    gp, gi, gx = r["nn"]
    G = sp.csc_matrix((gx, gi, gp), shape=(n, n))
    assert (np.asarray(G.sum(axis=1)).ravel() == k).all() and gp[-1] == n * k
    # idempotence: SNN recomputed from the returned index is identical
    g2 = sb.ComputeSNN(idx, 1 / 15, ctx=ctx)
    assert_csr_equal((g2.p, g2.i, g2.x), (p, i, x), "recomputed snn")
    # and it matches the oracle's SNN on the same index (bit-exact given identical kNN)
    assert_csr_equal((p, i, x), oracle.compute_snn(idx, 1 / 15), "snn vs oracle at 100k")


def test_dev_entry_points_and_row_sharding(ctx):
    """device-buffer ABI: kNN on a row shard + SNN rows of that shard == slices of the full result"""
    import torch
    from seurat_b200 import dist as sdist
    X, _ = gaussian_mixture(12000, 30, seed=2)
    k = 30
    want = oracle.find_neighbors(X, k=k, prune=1 / 15, method="kdtree")
    be = sdist.CudaBackend(0)
    ref = torch.from_numpy(X).cuda()
    parts = []
    for (b, e) in sdist.shard_bounds(len(X), 3):
        idx0, d = be.knn_shard(ref, b, e, k)
        assert (idx0.cpu().numpy() + 1 == want["nn_idx"][b:e]).all()
        assert (d.cpu().numpy() == want["nn_dist"][b:e]).all()
        parts.append(idx0)
    idx_full = torch.cat(parts)
    P, I, Xv = want["snn"]
    for (b, e) in sdist.shard_bounds(len(X), 3):
        g = be.graph_rows(idx_full, k, 1 / 15, True, b, e)
        p, i, x = (t.cpu().numpy() for t in g["snn"])
        assert (p == P[b:e + 1] - P[b]).all()
        assert (i == I[P[b]:P[e]]).all() and (x == Xv[P[b]:P[e]]).all()
        assert_csr_equal(tuple(t.cpu().numpy() for t in g["nn"]), want["nn"], "nn graph (every rank holds it all)")


@pytest.mark.parametrize("layout", ["C", "F"])
def test_resident_index_repeated_queries(ctx, layout):
    """NNHelper(cache_index=True) keeps the reference on the device; FindNeighbors(index=) / repeated
    queries against it give exactly what a fresh call gives (R/clustering.R:596-597, :1686-1688)."""
    ref, c = gaussian_mixture(6000, 20, seed=31, n_centres=5)
    q1, _ = gaussian_mixture(700, 20, seed=32, centres=c)
    q2, _ = gaussian_mixture(33, 20, seed=33, centres=c)
    R = sb.Embedding(np.asarray(ref, order=layout), cell_names(6000))
    first = sb.NNHelper(R, k=15, method="rann", cache_index=True, ctx=ctx)
    assert isinstance(first.alg_idx, _lib.Index) and first.alg_idx.rows == 6000 and first.alg_idx.dims == 20
    wi, wd = oracle.knn(ref, k=15, method="kdtree")
    assert (first.nn_idx == wi).all() and (first.nn_dist == wd).all()
    for q in (q1, q2):
        Q = sb.Embedding(np.asarray(q, order=layout), cell_names(q.shape[0]))
        nb = sb.FindNeighbors(R, query=Q, k_param=15, return_neighbor=True, index=first.alg_idx, verbose=False,
                              ctx=ctx)
        wi, wd = oracle.knn(ref, q, 15, method="kdtree")
        assert (nb.nn_idx == wi).all() and (nb.nn_dist == wd).all()
        assert nb.alg_idx is None and nb.cell_names == Q.rownames
    # raw ABI: self query through the index, k too large, foreign context
    idx, dist, st = first.alg_idx.knn(None, 15)
    assert (idx == first.nn_idx).all() and st["ms_h2d"] < 1.0
    with pytest.raises(sb.B200Error, match="Cannot find more nearest neighbours"):
        first.alg_idx.knn(None, 6001)
    with pytest.raises(ValueError, match="does not match"):
        sb.NNHelper(sb.Embedding(ref[:100], cell_names(100)), k=5, index=first.alg_idx, ctx=ctx)
    first.alg_idx.close()
    with pytest.raises(sb.B200Error):
        first.alg_idx.knn(None, 5)


def test_find_neighbors_into_caller_buffers(ctx):
    """nn Graph buffers handed to the count call are filled there (while the SNN stage runs);
    the result equals the plain two-phase call"""
    X, _ = gaussian_mixture(5000, 12, seed=41, n_centres=4)
    n, k = 5000, 10
    want = ctx.find_neighbors(X, k, 1 / 15, True, _lib.ROW_MAJOR)
    out = {"idx": np.empty((n, k), np.int32), "dist": np.empty((n, k), np.float64),
           "nn_p": np.empty(n + 1, np.int32), "nn_i": np.empty(n * k + 7, np.int32),
           "nn_x": np.empty(n * k + 7, np.float64)}
    got = ctx.find_neighbors(X, k, 1 / 15, True, _lib.ROW_MAJOR, out=out)
    assert got["nn"][0].base is out["nn_p"] or got["nn"][0] is out["nn_p"]
    for a, b in zip(want["nn"] + want["snn"] + (want["idx"], want["dist"]),
                    got["nn"] + got["snn"] + (got["idx"], got["dist"])):
        assert a.shape == b.shape and (a == b).all()
    only_nn = ctx.find_neighbors(X, k, 1 / 15, False, _lib.ROW_MAJOR, out=out)
    assert only_nn["snn"] is None and (only_nn["nn"][1] == want["nn"][1]).all()


@pytest.mark.parametrize("n,d,k", [(1, 3, 1), (2, 1, 2), (7, 1, 1), (33, 2, 33), (257, 5, 1), (3000, 61, 3), (2100, 1, 32)])
def test_knn_degenerate_shapes(ctx, n, d, k):
    """single points, one dimension, k = 1, k = n: both algorithms agree with the oracle"""
    rng = np.random.default_rng(n * 131 + d)
    X = rng.normal(size=(n, d))
    wi, wd = oracle.knn(X, k=k)
    for algo in ALGOS:
        if algo == _lib.KNN_TENSOR and n < 2048:
            continue
        idx, dist, _ = ctx.knn(X, None, k, _lib.ROW_MAJOR, algo)
        assert (idx == wi).all() and (dist == wd).all(), (n, d, k, algo)
    q = rng.normal(size=(1, d))                     # a single query row
    wi, wd = oracle.knn(X, q, k)
    idx, dist, _ = ctx.knn(X, q, k, _lib.ROW_MAJOR)
    assert (idx == wi).all() and (dist == wd).all()


@pytest.mark.parametrize("n,k", [(1, 1), (2, 1), (5, 5), (40, 1)])
def test_graphs_degenerate_shapes(ctx, n, k):
    rng = np.random.default_rng(n + k)
    nn = np.stack([rng.permutation(n)[:k] + 1 for _ in range(n)]).astype(np.int32)
    for prune in (0.0, 1 / 15, 1.0):
        want = oracle.compute_snn(nn, prune)
        p, i, x, _ = ctx.compute_snn(nn, prune, _lib.ROW_MAJOR)
        assert_csr_equal((p, i, x), want)
    wp, wi_, wx = oracle.nn_graph(nn, n)
    p, i, x, _ = ctx.nn_graph(nn, n, _lib.ROW_MAJOR)
    assert_csr_equal((p, i, x), (wp, wi_, wx))
