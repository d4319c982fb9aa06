"""CPU suite: host-side argument handling mirrored from FindNeighbors.default /
NNHelper / L2Norm (R/clustering.R:582-685, 1658-1690; R/utilities.R:2107-2122)."""
import numpy as np
import pytest

import oracle
import seurat_b200 as sb
from seurat_b200 import dist as sdist


def test_rownames_required():
    with pytest.raises(ValueError, match="Please provide rownames"):
        sb.FindNeighbors(np.zeros((10, 3)), k_param=3, verbose=False)


def test_invalid_method():
    X = sb.Embedding(np.zeros((10, 3)), [str(i) for i in range(10)])
    with pytest.raises(ValueError, match="Invalid method"):
        sb.NNHelper(X, k=3, method="nope")
    with pytest.raises(NotImplementedError):
        sb.NNHelper(X, k=3, method="annoy")


def test_k_too_large_message():
    X = sb.Embedding(np.random.default_rng(0).normal(size=(5, 3)), list("abcde"))
    with pytest.raises((ValueError, sb.B200Error), match="Cannot find more nearest neighbours|no CUDA"):
        sb.NNHelper(X, k=6, method="rann")


def test_non_finite_input_refused():
    v = np.random.default_rng(0).normal(size=(8, 3))
    v[2, 1] = np.nan
    X = sb.Embedding(v, [str(i) for i in range(8)])
    with pytest.raises((ValueError, sb.B200Error), match="NA/NaN/Inf|no CUDA"):
        sb.NNHelper(X, k=3, method="rann")


def test_l2norm_matches_reference_restatement():
    rng = np.random.default_rng(2)
    M = rng.normal(size=(40, 9)) * 10
    M[5] = 0.0
    got = sb.L2Norm(M)
    want = oracle.l2norm(M)
    assert (got == want).all()
    e = sb.L2Norm(sb.Embedding(M, [str(i) for i in range(40)]))
    assert isinstance(e, sb.Embedding) and (e.values == want).all()


def test_embedding_accepts_dataframe_and_fortran_order():
    import pandas as pd
    df = pd.DataFrame(np.arange(12.0).reshape(4, 3), index=["a", "b", "c", "d"])
    e = sb.Embedding(df)
    assert e.rownames == ["a", "b", "c", "d"] and e.shape == (4, 3)
    f = sb.Embedding(np.asfortranarray(np.arange(12.0).reshape(4, 3)), list("wxyz"))
    assert f.values.flags.f_contiguous


def test_shard_bounds():
    assert sdist.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sdist.shard_bounds(5, 4) == [(0, 2), (2, 4), (4, 5), (5, 5)]
    assert sdist.shard_bounds(8, 1) == [(0, 8)]
    for n in (1, 7, 100, 1_000_003):
        for w in (1, 2, 4, 8):
            b = sdist.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            # every shard except the tail has the common (padded) length
            per = b[0][1] - b[0][0]
            assert all(e - s <= per for s, e in b)


def test_index_argument_is_type_checked():
    X = sb.Embedding(np.random.default_rng(0).normal(size=(30, 4)), [str(i) for i in range(30)])
    with pytest.raises(TypeError, match="cache_index=True"):
        sb.NNHelper(X, k=3, method="rann", index=object())


def test_compute_snn_width_argument_check_needs_no_device():
    import seurat_b200 as sb
    g = sb.Graph(np.array([0, 1, 2], np.int32), np.array([0, 1], np.int32), np.array([1.0, 1.0]), (2, 2))
    with pytest.raises(ValueError, match=r"columns in the snn.graph \(2\)"):
        sb.ComputeSNNwidth(g, np.ones((2, 3)), 1, nearest_dist=[0.0, 0.0, 0.0])


def test_find_neighbors_resident_refuses_host_matrices():
    import seurat_b200 as sb
    with pytest.raises(TypeError, match="CUDA tensor"):
        sb.FindNeighborsResident(np.zeros((5, 2)), [f"c{i}" for i in range(5)])
