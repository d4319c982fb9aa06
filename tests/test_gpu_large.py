"""GPU suite, full-size configurations (-m gpu): the sizes the benchmark numbers are quoted on.

VERDICT r1 items 1b / "what's weak" 2: the grouped tensor search changes regime with n (number of
groups, item / tile geometry), so configs C, D and E of BASELINE.json and the inputs without coarse
cluster structure are checked here at full size, through the host-buffer C ABI:

  * kNN: >= 2000 sampled rows (plus the first and last rows) against the oracle's kd-tree search
    over the FULL reference set -- indices and distances bit-identical;
  * SNN: the whole graph (config C, shaped inputs) or a 100k-row block (config D) against the
    oracle's ComputeSNN restatement on the index the GPU produced -- (p, i, x) bit-identical;
  * size-independent properties (self first, ascending distances, symmetric SNN, unit diagonal).
"""
import numpy as np
import pytest

import oracle
import seurat_b200 as sb
from conftest import assert_csr_equal
from seurat_b200 import _lib
from seurat_b200.synthetic import CONFIG_E, CONFIGS, gaussian_mixture, shaped_embedding

pytestmark = pytest.mark.gpu

PRUNE = 1 / 15


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def sample_rows(n, m, seed=0):
    rows = np.random.default_rng(seed).choice(n, size=min(m, n), replace=False)
    return np.unique(np.concatenate([rows, np.arange(min(64, n)), np.arange(max(0, n - 64), n)]))


def check_knn_rows(X, Q, idx, dist, k, rows, what, method="kdtree"):
    # the kd-tree's order among exactly tied distances is its visiting order; inputs with exact
    # duplicates are checked against the brute-force scan (ties by lower index, like the library)
    wi, wd = oracle.knn(X, Q[rows], k, method=method)
    bad = np.nonzero((idx[rows] != wi).any(axis=1) | (dist[rows] != wd).any(axis=1))[0]
    assert bad.size == 0, f"{what}: {bad.size} of {rows.size} sampled rows differ from the oracle, first {rows[bad[:5]]}"


def check_knn_properties(idx, dist, n_ref, self_query):
    assert idx.min() >= 1 and idx.max() <= n_ref
    assert (np.diff(dist, axis=1) >= 0).all(), "distances must ascend"
    srt = np.sort(idx, axis=1)
    assert (srt[:, 1:] != srt[:, :-1]).all(), "a row lists a neighbour twice"
    if self_query:
        assert (dist[:, 0] == 0).all()
        assert (idx[:, 0] == np.arange(1, idx.shape[0] + 1)).all()


def check_snn_properties(p, i, x, n, k, rows):
    import scipy.sparse as sp
    S = sp.csc_matrix((x, i, p), shape=(n, n))
    assert (S != S.T).nnz == 0, "SNN must be symmetric"
    assert (S.diagonal() == 1.0).all()
    allowed = np.array([s / (k + (k - s)) for s in range(1, k + 1)])
    assert np.isin(x, allowed[allowed >= PRUNE]).all()
    assert all((np.diff(i[p[c]:p[c + 1]]) > 0).all() for c in rows)


def test_config_c_full_size(ctx):
    """config C: 1M x 50, k = 20, prune 1/15 -- the configuration the headline metric is quoted on"""
    cfg = CONFIGS["C"]
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = gaussian_mixture(n, d, cfg["seed"])
    r = ctx.find_neighbors(X, k, PRUNE, True, _lib.ROW_MAJOR)
    idx, dist = r["idx"], r["dist"]
    check_knn_properties(idx, dist, n, True)
    rows = sample_rows(n, 2000)
    check_knn_rows(X, X, idx, dist, k, rows, "config C kNN")
    assert r["stats"]["knn_tiles_visited"] < r["stats"]["knn_tiles_total"]    # the grouped search ran
    p, i, x = r["snn"]
    assert_csr_equal((p, i, x), oracle.compute_snn(idx, PRUNE), "config C SNN vs oracle on the same index")
    check_snn_properties(p, i, x, n, k, rows[:200])
    gp, gi, gx = r["nn"]
    assert_csr_equal((gp, gi, gx), oracle.nn_graph(idx), "config C nn Graph")


def test_config_d_full_size(ctx):
    """config D: 4M x 30, k = 30 (on one GPU here; the sharded run is in test_gpu_multi.py)"""
    cfg = CONFIGS["D"]
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = gaussian_mixture(n, d, cfg["seed"])
    r = ctx.find_neighbors(X, k, PRUNE, True, _lib.ROW_MAJOR)
    idx, dist = r["idx"], r["dist"]
    check_knn_properties(idx, dist, n, True)
    rows = sample_rows(n, 2000)
    check_knn_rows(X, X, idx, dist, k, rows, "config D kNN")
    p, i, x = r["snn"]
    assert p[-1] == i.shape[0] == x.shape[0]
    for (b, e) in ((0, 60_000), (n - 40_000, n)):
        wp, wi, wx = oracle.compute_snn_rows(idx, PRUNE, b, e)
        assert (p[b:e + 1].astype(np.int64) - p[b] == wp).all(), "config D SNN block: p differs"
        assert (i[p[b]:p[e]] == wi).all() and (x[p[b]:p[e]].view(np.int64) == wx.view(np.int64)).all(), \
            "config D SNN block differs from the oracle"
    gp, gi, gx = r["nn"]
    assert gp[-1] == n * k and (gx == 1.0).all()


def test_config_e_query_path(ctx):
    """config E: 500k query cells against 2M reference cells x 50, k = 20 (Neighbor only)"""
    c = CONFIG_E
    R, cen = gaussian_mixture(c["nr"], c["d"], c["seed_ref"])
    Q, _ = gaussian_mixture(c["nq"], c["d"], c["seed_qry"], centres=cen)
    idx, dist, st = ctx.knn(R, Q, c["k"], _lib.ROW_MAJOR)
    check_knn_properties(idx, dist, c["nr"], False)
    check_knn_rows(R, Q, idx, dist, c["k"], sample_rows(c["nq"], 2000), "config E kNN")
    # the same queries through a resident index give the same answer
    ix = ctx.index(R, _lib.ROW_MAJOR)
    idx2, dist2, _ = ix.knn(Q[:50_000], c["k"])
    assert (idx2 == idx[:50_000]).all() and (dist2 == dist[:50_000]).all()
    ix.close()


def test_resident_index_keeps_prepared_reference(ctx):
    """b200nn_index (NNHelper(cache.index = TRUE)) in the grouped regime: the first query prepares
    the reference side (FP16 operands, grouping, balls) and later queries reuse it -- same answers as fresh
    calls, for other query sets, other k, the self query, and after unrelated calls on the same context"""
    R, cen = gaussian_mixture(120_000, 50, seed=61, n_centres=16)
    Q1, _ = gaussian_mixture(30_000, 50, seed=62, centres=cen)
    Q2, _ = gaussian_mixture(7_001, 50, seed=63, centres=cen)
    Q2[5] *= 40.0                                   # far outside the reference's range: still exact
    ix = ctx.index(R, _lib.ROW_MAJOR)
    stats = []
    # k = 50 needs three column splits: the prepared layout changes, the index re-prepares once and again for k = 31
    for q, k in ((Q1, 20), (Q2, 20), (Q1, 7), (None, 20), (Q2, 50), (Q2[:3000], 50), (Q2, 31)):
        gi, gd, st = ix.knn(q, k)
        stats.append(st)
        wi, wd, _ = ctx.knn(R, q, k, _lib.ROW_MAJOR)      # a fresh call (also disturbs the shared workspace)
        assert (gi == wi).all() and (gd == wd).all()
        rows = sample_rows(gi.shape[0], 300)
        qq = R if q is None else q
        oi, od = oracle.knn(R, qq[rows], k, method="kdtree")
        assert (gi[rows] == oi).all() and (gd[rows] == od).all()
    assert stats[1]["ms_prep"] < 0.6 * stats[0]["ms_prep"], [s["ms_prep"] for s in stats]
    bad = Q1[:100].copy()
    bad[3, 3] = np.inf
    with pytest.raises(sb.B200Error, match="NA/NaN/Inf"):
        ix.knn(bad, 5)
    gi, gd, _ = ix.knn(Q1[:100], 5)                 # the index survives the refusal
    wi, wd, _ = ctx.knn(R, Q1[:100], 5, _lib.ROW_MAJOR)
    assert (gi == wi).all() and (gd == wd).all()
    ix.close()


@pytest.mark.parametrize("kind", ["pca_like", "manifold10", "uniform", "dups"])
def test_shaped_inputs_250k(ctx, kind):
    """inputs without the benchmark's coarse clusters: still exact (and the pruning must not assume them)"""
    n, d, k = 250_000, 50, 20
    X = shaped_embedding(kind, n, d)
    r = ctx.find_neighbors(X, k, PRUNE, True, _lib.ROW_MAJOR)
    idx, dist = r["idx"], r["dist"]
    if kind != "dups":
        check_knn_properties(idx, dist, n, True)
    check_knn_rows(X, X, idx, dist, k, sample_rows(n, 1500), f"{kind} kNN", "brute" if kind == "dups" else "kdtree")
    assert_csr_equal(r["snn"], oracle.compute_snn(idx, PRUNE), f"{kind} SNN")


def test_non_finite_input_refused(ctx):
    """RANN::nn2 refuses NA/NaN/Inf (.C(NAOK = FALSE)); so does the C ABI, on both kNN paths"""
    rng = np.random.default_rng(3)
    for n, algo in ((300, _lib.KNN_EXACT), (5000, _lib.KNN_AUTO)):
        for bad in (np.nan, np.inf, -np.inf):
            X = rng.normal(size=(n, 8))
            X[n // 2, 3] = bad
            with pytest.raises(sb.B200Error, match="NA/NaN/Inf") as ei:
                ctx.knn(X, None, 5, _lib.ROW_MAJOR, algo)
            assert ei.value.code == _lib.ERR_INVALID
            Xg = rng.normal(size=(n, 8))
            with pytest.raises(sb.B200Error, match="NA/NaN/Inf"):
                ctx.knn(Xg, X[: n // 2 + 1], 5, _lib.ROW_MAJOR, algo)
        # and the context keeps working afterwards
        Xg = rng.normal(size=(n, 8))
        idx, dist, _ = ctx.knn(Xg, None, 5, _lib.ROW_MAJOR, algo)
        wi, wd = oracle.knn(Xg, k=5)
        assert (idx == wi).all() and (dist == wd).all()


@pytest.mark.parametrize("k,prune", [(30, 1 / 15), (32, 0.1), (33, 1 / 15), (10, 1 / 15), (20, 0.2), (8, 1 / 15)])
def test_snn_single_walk_kernel_shapes(ctx, k, prune):
    """the single-walk SNN kernel (k <= 32, 2 <= s_min <= 12) and its neighbours in the dispatch
    (k = 33 two-walk kernel, k = 8 with s_min = 1 plain kernel) against the oracle"""
    X, _ = gaussian_mixture(40_000, 12, seed=77 + k, n_centres=6)
    idx, _ = oracle.knn(X, k=k, method="kdtree")
    g = sb.ComputeSNN(idx, prune, ctx=ctx)
    assert_csr_equal((g.p, g.i, g.x), oracle.compute_snn(idx, prune), f"k={k} prune={prune}")
