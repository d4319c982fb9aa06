"""CPU suite: the oracle (oracle/oracle.c) against the committed golden fixtures and
independent scipy / sklearn restatements.  The fixtures were produced by
tests/golden/make_golden.py WITHOUT the oracle (pure Python + scipy)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import assert_csr_equal, unhex, unhex2


def test_survey_known_answers(fixtures):
    s = fixtures["survey_n6"]
    nn = np.array(s["nn_ranked"], dtype=np.int32)
    for key, prune in (("prune_0", 0.0), ("prune_1/15", 1 / 15), ("prune_0.21", 0.21)):
        want = (np.array(s[key]["p"]), np.array(s[key]["i"]), np.array(s[key]["x"], dtype=np.float64))
        assert_csr_equal(oracle.compute_snn(nn, prune), want, key)


def test_snn_edge_cases(fixtures):
    for name, c in fixtures["snn_cases"].items():
        nn = np.array(c["nn_ranked"], dtype=np.int32)
        prune = float.fromhex(c["prune"])
        want = (np.array(c["snn"]["p"]), np.array(c["snn"]["i"]), unhex(c["snn"]["x"]))
        assert_csr_equal(oracle.compute_snn(nn, prune), want, name)
        assert_csr_equal(oracle.compute_snn(nn, prune, nthreads=3), want, name + " (3 threads)")
        wantg = (np.array(c["nn"]["p"]), np.array(c["nn"]["i"]), unhex(c["nn"]["x"]))
        assert_csr_equal(oracle.nn_graph(nn), wantg, name + " nn graph")


def test_equality_with_cutoff_is_kept(fixtures):
    # snn.cpp:31 uses a strict `<`: k=8,s=1 / k=16,s=2 / k=24,s=3 give exactly 1/15 and stay
    for k, s in ((8, 1), (16, 2), (24, 3)):
        c = fixtures["snn_cases"][f"equality_kept_k{k}"]
        x = unhex(c["snn"]["x"])
        v = s / (k + (k - s))
        assert v == 1 / 15
        assert (x == v).any(), f"no entry with overlap {s} in the k={k} case"
        assert not (x < 1 / 15).any()


def test_pbmc_fixture(pbmc):
    for method in ("brute", "kdtree"):
        idx, dist = oracle.knn(pbmc["X"], k=pbmc["k"], method=method)
        assert (idx == pbmc["nn_idx"]).all(), method
        assert (dist == pbmc["nn_dist"]).all(), method  # bit-exact: same summation order
    assert (pbmc["nn_idx"][:, 0] == np.arange(1, 81)).all()  # self first
    assert_csr_equal(oracle.nn_graph(pbmc["nn_idx"]), pbmc["nn"], "pbmc nn graph")
    assert_csr_equal(oracle.compute_snn(pbmc["nn_idx"], pbmc["prune"]), pbmc["snn"], "pbmc snn")


def test_pbmc_query_path(fixtures, pbmc):
    q = fixtures["pbmc_query"]
    X = pbmc["X"]
    ref, qry = X[q["ref_rows"][0]:q["ref_rows"][1]], X[q["query_rows"][0]:q["query_rows"][1]]
    for method in ("brute", "kdtree"):
        idx, dist = oracle.knn(ref, qry, k=q["k"], method=method)
        assert (idx == np.array(q["nn_idx"])).all()
        assert (dist == unhex2(q["nn_dist"])).all()


def test_k_larger_than_points_errors():
    X = np.random.default_rng(0).normal(size=(5, 3))
    with pytest.raises(ValueError, match="Cannot find more nearest neighbours"):
        oracle.knn(X, k=6)


def test_kdtree_equals_brute_random():
    rng = np.random.default_rng(7)
    for n, d, k in ((500, 2, 5), (1500, 10, 20), (800, 30, 30), (64, 50, 63)):
        X = rng.normal(size=(n, d)) * rng.uniform(0.5, 3, size=d)
        Q = rng.normal(size=(200, d))
        ib, db = oracle.knn(X, Q, k, method="brute")
        it, dt = oracle.knn(X, Q, k, method="kdtree")
        assert (db == dt).all()
        assert (ib == it).all()  # continuous data: no ties


def test_ties_ordered_by_lower_index():
    X = np.zeros((10, 4))
    X[5:] = 1.0
    idx, dist = oracle.knn(X, k=6, method="brute")
    assert idx[0].tolist() == [1, 2, 3, 4, 5, 6]
    assert idx[7].tolist() == [6, 7, 8, 9, 10, 1]
    assert dist[0].tolist() == [0, 0, 0, 0, 0, 2.0]


def test_snn_against_scipy_random():
    rng = np.random.default_rng(3)
    X = rng.normal(size=(3000, 8))
    idx, _ = oracle.knn(X, k=15)
    n, k = idx.shape
    for prune in (0.0, 1 / 15, 0.3):
        A = sp.csr_matrix((np.ones(n * k), (np.repeat(np.arange(n), k), (idx - 1).ravel())), shape=(n, n))
        S = (A @ A.T).tocsr()
        S.data = S.data / (k + (k - S.data))
        S.data[S.data < prune] = 0
        S.eliminate_zeros()
        S.sort_indices()
        assert_csr_equal(oracle.compute_snn(idx, prune), (S.indptr, S.indices, S.data), f"prune={prune}")


def test_knn_against_sklearn():
    from sklearn.neighbors import NearestNeighbors
    rng = np.random.default_rng(11)
    X = rng.normal(size=(2000, 20))
    idx, dist = oracle.knn(X, k=10)
    d2, i2 = NearestNeighbors(n_neighbors=10, algorithm="brute").fit(X).kneighbors(X)
    assert (i2 + 1 == idx).all()
    assert np.allclose(d2, dist, rtol=1e-6, atol=1e-5)  # sklearn uses the less accurate |x|^2+|y|^2-2xy


def test_l2norm():
    rng = np.random.default_rng(5)
    M = rng.normal(size=(50, 7))
    M[3] = 0.0  # 0/0 -> non-finite -> 0
    out = oracle.l2norm(M)
    assert (out[3] == 0).all()
    nrm = np.sqrt((out ** 2).sum(axis=1))
    assert np.allclose(np.delete(nrm, 3), 1.0)


def test_oracle_selftest_under_asan_ubsan():
    """the checker is itself checked: every oracle entry point on small inputs under
    AddressSanitizer + UBSan (leaks included), plus the SURVEY 8c vector from C"""
    import os
    import subprocess
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    b = subprocess.run(["make", "-C", root, "selftest_asan"], capture_output=True, text=True, timeout=300)
    if b.returncode != 0 and "sanitize" in (b.stderr + b.stdout):
        pytest.skip("compiler without sanitizer runtime")
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([os.path.join(root, "selftest_asan")], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0 and "oracle selftest ok" in r.stdout, (r.stdout + r.stderr)[-3000:]


def test_knn_against_scipy_ckdtree_20k():
    """a third, independent exact search (scipy.spatial.cKDTree: its own kd-tree and distance loop) on the
    benchmark's kind of input at 20k x 50 and on a query != reference case: same neighbours wherever the
    distances are distinct, distances equal to the last few bits"""
    from scipy.spatial import cKDTree
    from seurat_b200.synthetic import gaussian_mixture
    X, cen = gaussian_mixture(20_000, 50, seed=5, n_centres=12)
    Q, _ = gaussian_mixture(3_000, 50, seed=6, centres=cen)
    for qry in (None, Q):
        idx, dist = oracle.knn(X, qry, 20, method="kdtree")
        d2, i2 = cKDTree(X, leafsize=8).query(X if qry is None else qry, k=20)
        assert np.allclose(d2, dist, rtol=1e-12, atol=0)
        gap_ok = np.ones_like(dist, dtype=bool)           # positions whose neighbours in the list are not tied
        gap_ok[:, 1:] &= (dist[:, 1:] - dist[:, :-1]) > 1e-9 * dist[:, 1:]
        gap_ok[:, :-1] &= (dist[:, 1:] - dist[:, :-1]) > 1e-9 * dist[:, 1:]
        assert (idx[gap_ok] == i2[gap_ok] + 1).all()
        assert gap_ok.mean() > 0.999
