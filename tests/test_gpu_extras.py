"""GPU suite (-m gpu): SNN side consumers through the C ABI against the oracle."""
import numpy as np
import pytest

import oracle
import seurat_b200 as sb
from seurat_b200 import _lib
from seurat_b200.synthetic import gaussian_mixture

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("order", ["C", "F"])
def test_fast_dist_bit_exact(ctx, order):
    rng = np.random.default_rng(2)
    x = np.asarray(rng.normal(size=(3000, 30)), order=order)
    y = np.asarray(rng.normal(size=(5000, 30)), order=order)
    lists = [rng.integers(1, 5001, rng.integers(0, 40)) for _ in range(3000)]
    want = oracle.fast_dist(x, y, lists)
    got = sb.fast_dist(x, y, lists, ctx=ctx)
    assert len(got) == 3000 and all((a == b).all() for a, b in zip(got, want))
    assert sb.fast_dist(x, y, lists[:10], ctx=ctx) == []                 # length mismatch: empty list
    with pytest.raises(sb.B200Error):
        sb.fast_dist(x[:2], y, [np.array([1, 5001]), np.array([2])], ctx=ctx)
    # the kNN distances themselves are such distances
    X, _ = gaussian_mixture(4000, 20, seed=3, n_centres=4)
    nb = sb.FindNeighbors(sb.Embedding(X, [f"c{i}" for i in range(4000)]), k_param=10, return_neighbor=True,
                          verbose=False, ctx=ctx)
    d = sb.fast_dist(X, X, list(nb.nn_idx), ctx=ctx)
    assert (np.stack(d) == nb.nn_dist).all()


@pytest.mark.parametrize("order", ["C", "F"])
def test_snn_smallest_nonzero_dist(ctx, order):
    X, _ = gaussian_mixture(6000, 15, seed=4, n_centres=5)
    X[11] = X[10]
    r = oracle.find_neighbors(X, k=20, prune=1 / 15, method="kdtree")
    p, i, x = r["snn"]
    g = sb.Graph(p, i, x, (6000, 6000))
    nearest = r["nn_dist"][:, 1].copy()
    nearest[::5] = 0.0
    Xo = np.asarray(X, order=order)
    for n_small in (1, 20, 100000):
        want = oracle.snn_smallest_nonzero_dist(p, i, x, X, n_small, nearest)
        got = sb.SNN_SmallestNonzero_Dist(g, Xo, n_small, nearest, ctx=ctx)
        assert (got == want).all()        # same summation order as the oracle: bit-identical


def test_direct_snn_to_file(ctx, tmp_path):
    X, _ = gaussian_mixture(5000, 10, seed=6, n_centres=3)
    r = oracle.find_neighbors(X, k=15, prune=1 / 15)
    f1, f2 = str(tmp_path / "gpu.txt"), str(tmp_path / "oracle.txt")
    g = sb.DirectSNNToFile(r["nn_idx"], 1 / 15, False, f1, ctx=ctx)
    oracle.write_edge_file(*r["snn"], f2)
    assert open(f1, "rb").read() == open(f2, "rb").read()
    assert (g.p == r["snn"][0]).all() and (g.x == r["snn"][2]).all()


def test_r_level_callers_of_the_side_consumers(ctx):
    """ComputeSNNwidth (R/clustering.R:1027-1049) and NNdist (R/clustering.R:1620-1644)"""
    X, _ = gaussian_mixture(3000, 12, seed=9, n_centres=4)
    r = oracle.find_neighbors(X, k=15, prune=0.0, method="kdtree")
    p, i, x = r["snn"]
    g = sb.Graph(p, i, x, (3000, 3000))
    Xn = sb.L2Norm(X)
    want = oracle.snn_smallest_nonzero_dist(p, i, x, Xn, 15, np.zeros(3000))
    assert (sb.ComputeSNNwidth(g, X, 15, ctx=ctx) == want).all()
    nearest = r["nn_dist"][:, 1]
    want = oracle.snn_smallest_nonzero_dist(p, i, x, X, 15, nearest)
    assert (sb.ComputeSNNwidth(g, X, 15, l2_norm=False, nearest_dist=nearest, ctx=ctx) == want).all()
    with pytest.raises(ValueError, match="as many elements as there are columns"):
        sb.ComputeSNNwidth(g, X, 15, nearest_dist=np.zeros(7), ctx=ctx)
    d = sb.NNdist(r["nn_idx"], X, ctx=ctx)                     # matrix of indices -> one row per cell
    assert (np.stack(d) == r["nn_dist"]).all()
    d0 = sb.NNdist(list(r["nn_idx"][:50]), X, query_embeddings=X[:50], nearest_dist=nearest[:50], ctx=ctx)
    w0 = np.maximum(r["nn_dist"][:50] - nearest[:50, None], 0.0)
    assert (np.stack(d0) == w0).all() and (np.stack(d0)[:, :2] == 0).all()


def test_find_neighbors_on_a_device_resident_embedding(ctx):
    """an embedding that already lives in HBM goes through dims slicing and L2Norm on the device
    (b200nn_dev_embedding_prepare) and through the dev_* stages -- same Neighbor / Graphs as the host path"""
    import torch
    n, D = 6000, 24
    X, _ = gaussian_mixture(n, D, seed=41, n_centres=6)
    X[17] = 0.0                                             # a zero row: L2Norm gives zeros, not NaN
    names = [f"c{i}" for i in range(n)]
    dims = [0, 1, 2, 3, 4, 5, 6, 7, 9, 11, 12, 20]          # Embeddings(object[[reduction]])[, dims]
    row_major = torch.from_numpy(X).cuda()
    col_major = torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()     # R's layout: the transpose is contiguous
    assert row_major.is_contiguous() and not col_major.is_contiguous() and col_major.t().is_contiguous()
    for l2 in (False, True):
        sub = X[:, dims]
        want_mat = oracle.l2norm(sub) if l2 else sub
        for t in (row_major, col_major):
            work = torch.empty((n, len(dims)), dtype=torch.float64, device="cuda")
            ctx.dev_embedding_prepare(t, n, D, _lib.ROW_MAJOR if t.is_contiguous() else _lib.COL_MAJOR, dims, l2, work)
            assert (work.cpu().numpy() == want_mat).all(), "device slice / L2Norm must equal R's arithmetic"
        assert (sb.L2Norm(sub) == oracle.l2norm(sub)).all()
        host = sb.FindNeighbors(sb.Embedding(sub, names), k_param=15, l2_norm=l2, verbose=False, ctx=ctx)
        nb_h = sb.FindNeighbors(sb.Embedding(sub, names), k_param=15, l2_norm=l2, return_neighbor=True,
                                verbose=False, ctx=ctx)
        for t in (row_major, col_major):
            dev = sb.FindNeighborsResident(t, names, dims=dims, k_param=15, l2_norm=l2, verbose=False, ctx=ctx)
            for name in ("nn", "snn"):
                a, b = dev[name], host[name]
                assert (a.p == b.p).all() and (a.i == b.i).all() and (a.x == b.x).all(), (name, l2)
                assert a.rownames == names and a.Dim == (n, n)
            nb_d = sb.FindNeighborsResident(t, names, dims=dims, k_param=15, l2_norm=l2, return_neighbor=True,
                                            verbose=False, ctx=ctx)
            assert (nb_d.nn_idx == nb_h.nn_idx).all() and (nb_d.nn_dist == nb_h.nn_dist).all()
    # every component of a row-major tensor, no normalisation: used in place; nn Graph only
    g = sb.FindNeighborsResident(row_major, names, k_param=10, compute_SNN=False, verbose=False, ctx=ctx)
    h = sb.FindNeighbors(sb.Embedding(X, names), k_param=10, compute_SNN=False, verbose=False, ctx=ctx)
    assert set(g) == {"nn"} and (g["nn"].i == h["nn"].i).all() and (g["nn"].p == h["nn"].p).all()
    with pytest.raises(sb.B200Error, match="subscript out of bounds"):
        sb.FindNeighborsResident(row_major, names, dims=[0, D], verbose=False, ctx=ctx)
    with pytest.raises(TypeError):
        sb.FindNeighborsResident(X, names, ctx=ctx)
