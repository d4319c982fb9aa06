"""GPU suite: the tensor-core kNN path (tcgen05 candidates + FP64 re-rank + certificate)."""
import numpy as np
import pytest
import torch

import oracle
import seurat_b200 as sb
from seurat_b200 import _lib
from seurat_b200.synthetic import gaussian_mixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def fp16_operands(X, scale, mu):
    y = scale * (X - mu[None, :X.shape[1]])
    yh = y.astype(np.float16).astype(np.float64)     # round-to-nearest-even, like __double2half
    return y, yh


@pytest.mark.parametrize("n,d", [(4096, 50), (3000, 30), (2500, 61), (5000, 7)])
def test_raw_keys_match_fp16_model(ctx, n, d):
    """accumulator == |yh_r|^2 - 2 yh_q.yh_r for the first 256 queries x 128 references:
    validates TMA swizzle, UMMA descriptors, TMEM addressing and the norm-column trick"""
    X, _ = gaussian_mixture(n, d, seed=d, n_centres=8)
    ref = torch.from_numpy(X).cuda()
    keys = torch.full((256, 128), float("nan"), dtype=torch.float32, device="cuda")
    cand = torch.empty((n, 32), dtype=torch.int32, device="cuda")
    tau = torch.empty(n, dtype=torch.float32, device="cuda")
    sm = np.zeros(65)
    torch.cuda.synchronize()
    ctx.dev_knn_candidates(ref, n, ref, n, d, 32, cand, tau, keys, sm, 1)
    scale, mu = sm[0], sm[1:]
    assert scale > 0 and np.log2(scale) == np.round(np.log2(scale))
    y, yh = fp16_operands(X, scale, mu)
    assert (yh ** 2).sum(1).max() <= 33000
    want = (yh[:128] ** 2).sum(1)[None, :] - 2.0 * yh[:256] @ yh[:128].T
    got = keys.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    M = (yh ** 2).sum(1).max() * 3
    err = np.abs(got - want).max()
    print(f"n={n} d={d} scale={scale} max key err={err:.3e} (M 2^-24 = {M * 2**-24:.3e})")
    assert err <= 16 * M * 2.0 ** -24, "FP32 accumulation error far above the certificate's model"
    # candidates: every true neighbour of the FP16-rounded problem below tau is in the list
    full = (yh ** 2).sum(1)[None, :] - 2.0 * yh[:64] @ yh.T
    c = cand.cpu().numpy()[:64]
    t = tau.cpu().numpy()[:64].astype(np.float64)
    eta = 64 * M * 2.0 ** -24
    for q in range(64):
        must = np.nonzero(full[q] < t[q] - eta)[0]
        assert set(must) <= set(c[q]), f"query {q}: a key below tau is missing from the candidates"
        assert len(set(c[q])) == 32 and c[q].min() >= 0 and c[q].max() < n


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n,d,k", [(20000, 50, 20), (9000, 30, 30), (4097, 50, 5), (2048, 10, 32), (33000, 2, 10),
                                   (70000, 50, 20)])
def test_tensor_path_equals_oracle(ctx, n, d, k, mode):
    """mode 1: exhaustive scan of every reference tile; mode 2: grouped, tile-pruned search"""
    X, _ = gaussian_mixture(n, d, seed=n % 97, n_centres=16)
    wi, wd = oracle.knn(X, k=k, method="kdtree" if d > 4 else "brute")
    idx, dist, st = ctx.knn(X, None, k, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, mode)
    assert st["ms_knn_cand"] > 0
    assert (dist == wd).all()
    assert (idx == wi).all()
    print(f"n={n} d={d} k={k}: uncertified rows {st['n_uncertified']} / {n}")
    assert st["n_uncertified"] < 0.05 * n, "the certificate should pass for almost every row"


def test_tensor_path_query_ne_reference(ctx):
    ref, c = gaussian_mixture(30000, 50, seed=3, n_centres=8)
    qry, _ = gaussian_mixture(7001, 50, seed=4, centres=c)
    qry[:10] += 40.0            # far outside the reference cloud
    wi, wd = oracle.knn(ref, qry, 20, method="kdtree")
    for mode in (1, 2):
        idx, dist, st = ctx.knn(ref, qry, 20, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, mode)
        assert (idx == wi).all() and (dist == wd).all(), f"scan mode {mode}"


def test_tensor_path_ties_and_duplicates(ctx):
    """exact duplicates and lattice ties defeat the certificate; the FP64 fallback must resolve them"""
    rng = np.random.default_rng(1)
    base = rng.integers(0, 4, size=(3000, 6)).astype(np.float64)
    X = np.vstack([base, base[:500]])
    wi, wd = oracle.knn(X, k=10)
    for mode in (1, 2):
        idx, dist, st = ctx.knn(X, None, 10, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, mode)
        assert (dist == wd).all() and (idx == wi).all(), f"scan mode {mode}"
        assert st["n_uncertified"] > 0
    Z = np.ones((2500, 20)) * 3.25
    idx, dist, st = ctx.knn(Z, None, 5, _lib.ROW_MAJOR, _lib.KNN_TENSOR)
    assert (idx == np.tile(np.arange(1, 6), (2500, 1))).all() and (dist == 0).all()


def test_tensor_path_badly_scaled_data(ctx):
    """huge offsets, tiny spreads and mixed magnitudes: results stay exact (certificate or fallback)"""
    rng = np.random.default_rng(2)
    X = rng.normal(size=(6000, 20))
    X[:, 0] *= 1e4
    X[:, 1] *= 1e-4
    X += 1e6
    wi, wd = oracle.knn(X, k=15, method="kdtree")
    for mode in (1, 2):
        idx, dist, st = ctx.knn(X, None, 15, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, mode)
        assert (dist == wd).all() and (idx == wi).all(), f"scan mode {mode}"
        print("badly scaled: mode", mode, "uncertified", st["n_uncertified"])


def test_tensor_path_rejects_unsupported_shapes(ctx):
    X = np.random.default_rng(0).normal(size=(3000, 70))
    with pytest.raises(sb.B200Error):
        ctx.knn(X, None, 10, _lib.ROW_MAJOR, _lib.KNN_TENSOR)      # d > 61
    idx, dist, st = ctx.knn(X, None, 10, _lib.ROW_MAJOR, _lib.KNN_AUTO)   # auto falls back to the exact scan
    assert st["ms_knn_cand"] == 0
    wi, wd = oracle.knn(X, k=10)
    assert (idx == wi).all() and (dist == wd).all()


@pytest.mark.parametrize("kind", ["uniform", "unbalanced", "duplicates", "line"])
def test_grouped_search_on_awkward_data(ctx, kind):
    """the tile-pruned search must stay exact when grouping is useless or lopsided"""
    rng = np.random.default_rng(5)
    n, d, k = 40000, 12, 10
    if kind == "uniform":            # no cluster structure at all: nothing can be skipped
        X = rng.uniform(-1, 1, size=(n, d))
    elif kind == "unbalanced":       # one huge blob, a few tiny far-away ones (some groups nearly empty)
        X = rng.normal(size=(n, d))
        X[:50] += 200.0
        X[50:60] -= 300.0
    elif kind == "duplicates":       # every point three times: ties at every rank boundary
        base = rng.normal(size=(n // 4, d))
        X = np.vstack([base, base, base, rng.normal(size=(n - 3 * (n // 4), d))])
    else:                            # points on a line: extreme anisotropy
        t = rng.uniform(0, 1000, size=n)
        X = np.outer(t, np.ones(d)) + 1e-3 * rng.normal(size=(n, d))
    wi, wd = oracle.knn(X, k=k, method="kdtree" if kind != "duplicates" else "brute")
    idx, dist, st = ctx.knn(X, None, k, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, 2)
    assert (dist == wd).all(), kind
    assert (idx == wi).all(), kind
    print(kind, "uncertified", st["n_uncertified"], "tiles visited", st["knn_tiles_visited"], "of", st["knn_tiles_total"])


def test_grouped_search_few_queries_many_references(ctx):
    ref, c = gaussian_mixture(50000, 20, seed=8, n_centres=6)
    qry, _ = gaussian_mixture(37, 20, seed=9, centres=c)
    wi, wd = oracle.knn(ref, qry, 12, method="kdtree")
    idx, dist, st = ctx.knn(ref, qry, 12, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, 2)
    assert (idx == wi).all() and (dist == wd).all()


@pytest.mark.parametrize("mode", [1, 2])
def test_many_uncertified_rows_take_the_exact_fallback(ctx, mode):
    """a candidate set only slightly wider than k leaves many rows uncertified: they are redone by the
    FP64 scan (the group-pruned one after a grouped search) and the result is still the oracle's"""
    X, _ = gaussian_mixture(60000, 30, seed=12, n_centres=10)
    k = 20
    wi, wd = oracle.knn(X, k=k, method="kdtree")
    idx, dist, st = ctx.knn(X, None, k, _lib.ROW_MAJOR, _lib.KNN_TENSOR, k + 1, mode)
    assert st["n_uncertified"] > 100, st
    assert (idx == wi).all() and (dist == wd).all()
    # query != reference, few uncertified rows (the split-slice variant of the fallback)
    Q, _ = gaussian_mixture(300, 30, seed=13, n_centres=10)
    wi, wd = oracle.knn(X, Q, k, method="kdtree")
    idx, dist, st = ctx.knn(X, Q, k, _lib.ROW_MAJOR, _lib.KNN_TENSOR, k + 2, mode)
    assert (idx == wi).all() and (dist == wd).all()
    print("mode", mode, "uncertified of 300:", st["n_uncertified"])


# ---------------------------------------------------------------- k > 32: column splits
@pytest.mark.parametrize("n,d,k", [(33000, 50, 33), (60000, 30, 50), (40000, 50, 64), (50000, 20, 100),
                                   (45000, 10, 200), (36000, 40, 256)])
def test_large_k_runs_per_column_split(ctx, n, d, k):
    """k > 32 (k' > 48) no longer fits one shared-memory list per row: the grouped search runs once per
    column split of the reference set and the re-rank sees the union of the splits' candidates.
    FindNeighbors takes any k.param (R/clustering.R:586); WNN asks for knn.range = 200 (R/clustering.R:55)"""
    X, _ = gaussian_mixture(n, d, seed=k, n_centres=12)
    rows = np.sort(np.random.default_rng(k).choice(n, 2500, replace=False))
    wi, wd = oracle.knn(X, X[rows], k, method="kdtree")
    idx, dist, st = ctx.knn(X, None, k, _lib.ROW_MAJOR, _lib.KNN_TENSOR)
    assert st["ms_knn_cand"] > 0
    assert (dist[rows] == wd).all()
    assert (idx[rows] == wi).all()
    print(f"n={n} d={d} k={k}: uncertified rows {st['n_uncertified']} / {n}, tiles visited "
          f"{st['knn_tiles_visited']} of {st['knn_tiles_total']}")
    assert st["n_uncertified"] < 0.05 * n, "the certificate should pass for almost every row"
    # every row is a valid neighbour list: self first (distance 0), ascending distances, no repeats
    assert (dist[:, 0] == 0).all() and (np.diff(dist, axis=1) >= 0).all()
    srt = np.sort(idx, axis=1)
    assert (srt[:, 1:] != srt[:, :-1]).all() and srt.min() >= 1 and srt.max() <= n


def test_large_k_query_path_shares_and_small_sets(ctx):
    ref, c = gaussian_mixture(48000, 30, seed=31, n_centres=8)
    qry, _ = gaussian_mixture(5000, 30, seed=32, centres=c)
    k = 50
    wi, wd = oracle.knn(ref, qry, k, method="kdtree")
    idx, dist, st = ctx.knn(ref, qry, k, _lib.ROW_MAJOR, _lib.KNN_TENSOR)
    assert st["ms_knn_cand"] > 0 and (idx == wi).all() and (dist == wd).all()
    # group-aligned shares of a self search (the multi-GPU path) with two column splits
    n = len(ref)
    rows = np.sort(np.random.default_rng(3).choice(n, 2000, replace=False))
    wi, wd = oracle.knn(ref, ref[rows], 40, method="kdtree")
    r = torch.from_numpy(ref).cuda()
    tot = torch.zeros((n, 40), dtype=torch.int32, device="cuda")
    for rank in range(3):
        i0 = torch.zeros((n, 40), dtype=torch.int32, device="cuda")
        dd = torch.full((n, 40), float("nan"), dtype=torch.float64, device="cuda")
        ctx.dev_knn(r, n, None, n, 30, 40, i0, dd, shard_rank=rank, shard_world=3)
        tot += i0
    assert (tot.cpu().numpy()[rows] + 1 == wi).all()
    # below the grouped regime one list per row is all there is: the tensor path refuses, auto takes the FP64 scan
    X = ref[:6000]
    with pytest.raises(sb.B200Error):
        ctx.knn(X, None, 50, _lib.ROW_MAJOR, _lib.KNN_TENSOR)
    idx, dist, st = ctx.knn(X, None, 50, _lib.ROW_MAJOR, _lib.KNN_AUTO)
    assert st["ms_knn_cand"] == 0
    wi, wd = oracle.knn(X, k=50, method="kdtree")
    assert (idx == wi).all() and (dist == wd).all()


def test_find_neighbors_k50_graphs(ctx):
    """the whole path at k.param = 50: kNN through two column splits, nn Graph, SNN (s_min = 7)"""
    X, _ = gaussian_mixture(40000, 30, seed=77, n_centres=10)
    got = ctx.find_neighbors(X, 50, 1 / 15, True, _lib.ROW_MAJOR)
    assert got["stats"]["ms_knn_cand"] > 0
    want = oracle.find_neighbors(X, k=50, prune=1 / 15, method="kdtree")
    assert (got["idx"] == want["nn_idx"]).all() and (got["dist"] == want["nn_dist"]).all()
    for name in ("nn", "snn"):
        for a, b in zip(got[name], want[name]):
            assert a.shape == b.shape and (np.asarray(a) == np.asarray(b)).all(), name
