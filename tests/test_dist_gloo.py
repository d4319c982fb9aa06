"""CPU suite: the N>1 path (row sharding, padded all-gather of the index, per-shard
SNN rows, CSR concatenation) at world_size 2 and 3 over gloo.  The compute stages
are supplied by the ORACLE here (tests may use it); on GPUs the same host code
drives libb200nn through seurat_b200.dist.CudaBackend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from seurat_b200 import dist as sdist


class OracleBackend:
    def knn_shard(self, ref, q_begin, q_end, k, qry=None):
        src = ref if qry is None else qry
        if q_end == q_begin:
            return torch.empty((0, k), dtype=torch.int32), torch.empty((0, k), dtype=torch.float64)
        idx, d = oracle.knn(ref.numpy(), src[q_begin:q_end].numpy(), k)
        return torch.from_numpy(idx - 1), torch.from_numpy(d)

    def graph_rows(self, idx0_full, k, prune, compute_snn, row_begin, row_end):
        nn1 = idx0_full.numpy() + 1
        p, i, x = oracle.compute_snn(nn1, prune)
        lo, hi = p[row_begin], p[row_end]
        out = {"nn": tuple(torch.from_numpy(np.asarray(a)) for a in oracle.nn_graph(nn1))}
        out["snn"] = (torch.from_numpy(p[row_begin:row_end + 1] - lo), torch.from_numpy(i[lo:hi].copy()),
                      torch.from_numpy(x[lo:hi].copy()))
        return out


class OwnedOracleBackend(OracleBackend):
    """a backend that, like libb200nn's grouped search, chooses a NON-contiguous share of the rows"""

    def knn_owned(self, ref, rank, world, k):
        n = ref.shape[0]
        mine = np.nonzero((np.arange(n) * 7 + 3) % world == rank)[0]
        idx = torch.zeros((n, k), dtype=torch.int32)
        d = torch.full((n, k), float("nan"), dtype=torch.float64)
        if mine.size:
            wi, wd = oracle.knn(ref.numpy(), ref.numpy()[mine], k)
            idx[mine] = torch.from_numpy(wi - 1)
            d[mine] = torch.from_numpy(wd)
        return idx, d


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, d, k, prune, tmpdir, owned=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        X = torch.from_numpy(rng.normal(size=(n, d)))
        r = sdist.find_neighbors_sharded(X, k, prune, True, backend=OwnedOracleBackend() if owned else OracleBackend())
        b, e = r["rows"]
        if owned:   # the distances of the rows this rank computed, at their original positions
            wi, wd = oracle.knn(X.numpy(), k=k)
            own = r["owned"].numpy()
            assert own.sum() > 0 and (r["dist"].numpy()[own] == wd[own]).all()
        assert (b, e) == sdist.shard_bounds(n, world)[rank]
        full = sdist.gather_csr_rows(*r["snn"], dst=0)
        # kNN rows gathered for the check
        idx_full = r["idx0_full"]
        if rank == 0:
            p, i, x = full
            np.savez(os.path.join(tmpdir, "out.npz"), p=p.numpy(), i=i.numpy(), x=x.numpy(),
                     idx=idx_full.numpy(), nn_p=r["nn"][0].numpy())
        # query path: kNN only
        Q = torch.from_numpy(rng.normal(size=(n // 2 + 1, d)))
        rq = sdist.find_neighbors_sharded(X, k, prune, False, backend=OracleBackend(), qry=Q)
        qb, qe = rq["rows"]
        wi, wd = oracle.knn(X.numpy(), Q.numpy(), k)
        assert (rq["idx0"].numpy() + 1 == wi[qb:qe]).all() and (rq["dist"].numpy() == wd[qb:qe]).all()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,owned", [(2, 501, False), (3, 100, False), (2, 64, False), (2, 333, True),
                                            (3, 200, True)])
def test_sharded_pipeline_equals_single_process(tmp_path, world, n, owned):
    d, k, prune = 6, 10, 1 / 15
    mp.spawn(_worker, args=(world, _free_port(), n, d, k, prune, str(tmp_path), owned), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "out.npz"))
    X = np.random.default_rng(42).normal(size=(n, d))
    idx, _ = oracle.knn(X, k=k)
    assert (got["idx"] + 1 == idx).all()
    p, i, x = oracle.compute_snn(idx, prune)
    assert (got["p"] == p).all() and (got["i"] == i).all() and (got["x"] == x).all()
    assert (got["nn_p"] == oracle.nn_graph(idx)[0]).all()


def test_all_gather_index_single_rank_is_identity():
    t = torch.arange(12, dtype=torch.int32).view(4, 3)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        assert sdist.all_gather_index(t, 4) is t
    finally:
        dist.destroy_process_group()
