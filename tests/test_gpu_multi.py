"""GPU suite (-m gpu): the sharded path.

  * b200nn_multi_* (several devices from one process, what the R shim binds) against the
    single-device entry points, bit for bit -- on a single-GPU box the shards share device 0
    (the sharding, exchange and gathering logic is the same), on a multi-GPU box every visible
    device takes a shard;
  * the one-process-per-GPU path (seurat_b200/dist.py over torch.distributed + NCCL, as bench.py
    runs it under torchrun) against the single-GPU result when >= 2 devices are visible.
"""
import os

import numpy as np
import pytest

import oracle
import seurat_b200 as sb
from conftest import assert_csr_equal
from seurat_b200 import _lib
from seurat_b200.synthetic import CONFIGS, cell_names, gaussian_mixture

# a wedged multi-device call must end the run, not hold the box (pytest-timeout kills the process)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

PRUNE = 1 / 15


def shard_layouts():
    n = _lib.device_count()
    out = [[0, 0], [0, 0, 0]]                  # shards sharing device 0
    if n >= 2:
        out.append(list(range(min(n, 8))))     # one shard per visible device
        if n >= 3:
            out.append([1, 0])                 # any order / subset of devices
    return out


@pytest.fixture(scope="module")
def single():
    c = sb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("devices", shard_layouts(), ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("order", ["C", "F"])
def test_multi_matches_single_device(single, devices, order):
    X, _ = gaussian_mixture(40_000, 50, seed=11, n_centres=12)
    X = np.asarray(X, order=order)
    layout = _lib.ROW_MAJOR if order == "C" else _lib.COL_MAJOR
    k = 20
    want = single.find_neighbors(X, k, PRUNE, True, layout)
    m = sb.MultiContext(devices)
    assert m.size == len(devices)
    got = m.find_neighbors(X, k, PRUNE, True, layout)
    assert (got["idx"] == want["idx"]).all() and (got["dist"] == want["dist"]).all()
    assert_csr_equal(got["nn"], want["nn"], "nn Graph")
    assert_csr_equal(got["snn"], want["snn"], "SNN gathered from the shards")
    assert got["stats"]["nnz_snn"] == want["stats"]["nnz_snn"]
    # nn Graph only, caller-provided (early) buffers
    out = {"nn_p": np.empty(len(X) + 1, np.int32), "nn_i": np.empty(len(X) * k, np.int32),
           "nn_x": np.empty(len(X) * k, np.float64)}
    only = m.find_neighbors(X, k, PRUNE, False, layout, out=out)
    assert only["snn"] is None
    assert_csr_equal(only["nn"], want["nn"], "nn Graph into caller buffers")
    # the Neighbor-only calls: self and query != reference
    idx, dist, _ = m.knn(X, None, k, layout)
    assert (idx == want["idx"]).all() and (dist == want["dist"]).all()
    Q, _ = gaussian_mixture(777, 50, seed=12, n_centres=12)
    Q = np.asarray(Q, order=order)
    wi, wd, _ = single.knn(X, Q, 7, layout)
    gi, gd, _ = m.knn(X, Q, 7, layout)
    assert (gi == wi).all() and (gd == wd).all()
    m.close()


def test_multi_through_the_reference_interface(single):
    """FindNeighbors(..., ctx = MultiContext) is the call the R shim makes with several GPUs visible"""
    X, _ = gaussian_mixture(9_000, 20, seed=5, n_centres=5)
    E = sb.Embedding(X, cell_names(len(X)))
    m = sb.MultiContext([0, 0] if _lib.device_count() < 2 else [0, 1])
    g = sb.FindNeighbors(E, k_param=15, verbose=False, ctx=m)
    w = oracle.find_neighbors(X, k=15, prune=PRUNE, method="kdtree")
    assert_csr_equal((g["snn"].p, g["snn"].i, g["snn"].x), w["snn"], "snn")
    assert_csr_equal((g["nn"].p, g["nn"].i, g["nn"].x), w["nn"], "nn")
    nb = sb.FindNeighbors(E, k_param=15, return_neighbor=True, verbose=False, ctx=m)
    assert (nb.nn_idx == w["nn_idx"]).all() and (nb.nn_dist == w["nn_dist"]).all()
    m.close()


def test_multi_edge_cases():
    m = sb.MultiContext([0, 0, 0, 0])
    # fewer rows than shards, tiny exact-path problems, k = n
    for n, d, k in ((3, 2, 2), (5, 4, 5), (130, 3, 9)):
        X = np.random.default_rng(n).normal(size=(n, d))
        got = m.find_neighbors(X, k, PRUNE, True, _lib.ROW_MAJOR)
        w = oracle.find_neighbors(X, k=k, prune=PRUNE)
        assert (got["idx"] == w["nn_idx"]).all() and (got["dist"] == w["nn_dist"]).all()
        assert_csr_equal(got["snn"], w["snn"], f"n={n}")
        assert_csr_equal(got["nn"], w["nn"], f"n={n} nn")
    with pytest.raises(sb.B200Error, match="Cannot find more nearest neighbours"):
        m.knn(np.zeros((4, 2)), None, 5, _lib.ROW_MAJOR)
    X = np.random.default_rng(0).normal(size=(4000, 6))
    X[3999, 0] = np.nan                       # only the last shard sees it: every shard must stop
    with pytest.raises(sb.B200Error, match="NA/NaN/Inf"):
        m.find_neighbors(X, 5, PRUNE, True, _lib.ROW_MAJOR)
    X[3999, 0] = 0.0
    got = m.find_neighbors(X, 5, PRUNE, True, _lib.ROW_MAJOR)   # and the context keeps working
    assert_csr_equal(got["snn"], oracle.find_neighbors(X, k=5, prune=PRUNE)["snn"], "after an error")
    with pytest.raises(sb.B200Error):
        sb.MultiContext([0, 99])
    m.close()


@pytest.mark.skipif(_lib.device_count() < 2, reason="needs >= 2 GPUs")
def test_multi_config_c_all_devices(single):
    """config C (1M x 50, k = 20) sharded over every visible device == the single-device result"""
    cfg = CONFIGS["C"]
    X, _ = gaussian_mixture(cfg["n"], cfg["d"], cfg["seed"])
    want = single.find_neighbors(X, cfg["k"], PRUNE, True, _lib.ROW_MAJOR)
    m = sb.MultiContext(None)
    got = m.find_neighbors(X, cfg["k"], PRUNE, True, _lib.ROW_MAJOR)
    assert (got["idx"] == want["idx"]).all() and (got["dist"] == want["dist"]).all()
    assert_csr_equal(got["snn"], want["snn"], "config C SNN, all devices")
    assert_csr_equal(got["nn"], want["nn"], "config C nn Graph, all devices")
    m.close()


@pytest.mark.skipif(_lib.device_count() < 4, reason="needs >= 4 GPUs")
def test_multi_config_d_all_devices():
    """config D (4M x 30, k = 30) on every visible device: sampled kNN rows + SNN blocks vs the oracle"""
    cfg = CONFIGS["D"]
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = gaussian_mixture(n, d, cfg["seed"])
    m = sb.MultiContext(None)
    got = m.find_neighbors(X, k, PRUNE, True, _lib.ROW_MAJOR)
    m.close()
    idx, dist = got["idx"], got["dist"]
    rows = np.sort(np.random.default_rng(0).choice(n, 2000, replace=False))
    wi, wd = oracle.knn(X, X[rows], k, method="kdtree")
    assert (idx[rows] == wi).all() and (dist[rows] == wd).all()
    p, i, x = got["snn"]
    per = -(-n // len(m.devices))
    for b in (0, per - 20_000, n - 40_000):          # inside a shard, across a shard boundary, the tail
        e = b + 40_000
        wp, wi2, wx = oracle.compute_snn_rows(idx, PRUNE, b, e)
        assert (p[b:e + 1].astype(np.int64) - p[b] == wp).all()
        assert (i[p[b]:p[e]] == wi2).all() and (x[p[b]:p[e]].view(np.int64) == wx.view(np.int64)).all()


def test_dev_knn_shares_partition_the_rows(single):
    """b200nn_opts.shard_rank / shard_world on ONE device: the shares of a self search are disjoint, cover
    every row, and their zero-filled index buffers sum to the full index (what the NCCL all-reduce does)"""
    import torch
    n, d, k = 150_000, 30, 30
    X, _ = gaussian_mixture(n, d, seed=8)
    want_i, want_d, _ = single.knn(X, None, k, _lib.ROW_MAJOR)
    ref = torch.from_numpy(X).cuda()
    for world in (3, 8):
        tot = torch.zeros((n, k), dtype=torch.int32, device="cuda")
        owners = torch.zeros(n, dtype=torch.int32, device="cuda")
        for rank in range(world):
            idx = torch.zeros((n, k), dtype=torch.int32, device="cuda")
            dd = torch.full((n, k), float("nan"), dtype=torch.float64, device="cuda")
            st = single.dev_knn(ref, n, None, n, d, k, idx, dd, shard_rank=rank, shard_world=world)
            own = ~torch.isnan(dd[:, 0])
            owners += own.int()
            assert abs(int(own.sum()) - n / world) < 0.2 * n / world, "shares should be balanced"
            assert (dd[own].cpu().numpy() == want_d[own.cpu().numpy()]).all()
            tot += idx
        assert (owners == 1).all()
        assert (tot.cpu().numpy() + 1 == want_i).all()
    # tiny problem (FP64 scan): contiguous shares through the same option
    Xs = np.random.default_rng(1).normal(size=(500, 4))
    refs = torch.from_numpy(Xs).cuda()
    wi, wd = oracle.knn(Xs, k=6)
    tot = torch.zeros((500, 6), dtype=torch.int32, device="cuda")
    for rank in range(4):
        idx = torch.zeros((500, 6), dtype=torch.int32, device="cuda")
        dd = torch.full((500, 6), float("nan"), dtype=torch.float64, device="cuda")
        single.dev_knn(refs, 500, None, 500, 4, 6, idx, dd, shard_rank=rank, shard_world=4)
        tot += idx
    assert (tot.cpu().numpy() + 1 == wi).all()


# ---------------------------------------------------------------- one process per GPU (NCCL)
def _nccl_worker(rank, world, port, n, d, k, seed, ret):
    import torch
    import torch.distributed as dist
    from seurat_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    X, _ = gaussian_mixture(n, d, seed)
    be = sdist.CudaBackend(rank)
    ref = torch.from_numpy(X).cuda()
    out = sdist.find_neighbors_sharded(ref, k, PRUNE, True, backend=be)
    idx_full = out["idx0_full"].cpu().numpy()
    snn = sdist.gather_csr_rows(*out["snn"])
    ret[f"dist{rank}"] = out["dist"].cpu().numpy()          # full size, NaN rows belong to other ranks
    if rank == 0:
        ret["idx"] = idx_full
        ret["snn"] = tuple(t.numpy() for t in snn)
        ret["nn"] = tuple(t.cpu().numpy() for t in out["nn"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_lib.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_ranks_match_single_gpu(single):
    """the torchrun path of bench.py (one rank per GPU, NCCL all-gather of the index, CSR blocks gathered
    on rank 0) gives bit for bit what one GPU gives"""
    import torch.multiprocessing as mp
    world = min(_lib.device_count(), 4)
    n, d, k = 120_000, 50, 20
    X, _ = gaussian_mixture(n, d, 21)
    want = single.find_neighbors(X, k, PRUNE, True, _lib.ROW_MAJOR)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_nccl_worker, args=(world, 29000 + os.getpid() % 1000, n, d, k, 21, ret), nprocs=world, join=True)
        assert (ret["idx"] + 1 == want["idx"]).all()
        owners = np.zeros(n, dtype=np.int32)
        for r in range(world):                               # the shares partition the rows
            dr = ret[f"dist{r}"]
            own = ~np.isnan(dr[:, 0])
            owners += own
            assert (dr[own] == want["dist"][own]).all()
        assert (owners == 1).all()
        p, i, x = ret["snn"]
        assert_csr_equal((p, i, x), want["snn"], "SNN gathered over NCCL")
        assert_csr_equal(ret["nn"], want["nn"], "nn Graph (rank 0)")
