"""Runs the larger BASELINE.json configs once on one GPU and checks size-independent properties.

    python tools/run_configs.py D        # 4M x 30, k = 30 (kNN + nn Graph + SNN)
    python tools/run_configs.py E        # 500k queries vs 2M references x 50, k = 20 (Neighbor only)
    python tools/run_configs.py D --n 1000000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker only)
import seurat_b200 as sb  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=["D", "E"])
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--multi", type=int, default=0, help="shard over this many devices through b200nn_multi_* (host buffers)")
a = ap.parse_args()
out = {"config": a.config}
if a.multi:
    # the single-process multi-device C ABI, end to end from (pageable) host buffers
    from seurat_b200 import _lib
    m = sb.MultiContext(list(range(a.multi)))
    out["devices"] = a.multi
    if a.config == "D":
        n, d, k = a.n or 4_000_000, 30, 30
        X, _ = gaussian_mixture(n, d, 2)
        for r in range(a.reps):
            t0 = time.perf_counter()
            g = m.find_neighbors(X, k, 1 / 15, True, _lib.ROW_MAJOR)
            t1 = time.perf_counter()
        out.update(n=n, d=d, k=k, e2e_ms=(t1 - t0) * 1e3, cells_per_s_e2e=n / (t1 - t0), stats=g["stats"])
        rows = np.random.default_rng(0).choice(n, 400, replace=False)
        wi, wd = oracle.knn(X, X[rows], k, method="kdtree")
        out["knn_sample_exact"] = bool((g["idx"][rows] == wi).all() and (g["dist"][rows] == wd).all())
        p, i, x = g["snn"]
        per = -(-n // a.multi)
        b = max(0, per - 10000)
        wp, wi2, wx = oracle.compute_snn_rows(g["idx"], 1 / 15, b, b + 20000)
        out["snn_block_across_shards_exact"] = bool((p[b:b + 20001].astype(np.int64) - p[b] == wp).all()
                                                    and (i[p[b]:p[b + 20000]] == wi2).all()
                                                    and (x[p[b]:p[b + 20000]] == wx).all())
        out["nnz_per_row"] = float(p[-1]) / n
    else:
        nr, nq, d, k = a.n or 2_000_000, (a.n // 4 if a.n else 500_000), 50, 20
        R, c = gaussian_mixture(nr, d, 3)
        Q, _ = gaussian_mixture(nq, d, 4, centres=c)
        for r in range(a.reps):
            t0 = time.perf_counter()
            idx, dist, st = m.knn(R, Q, k, _lib.ROW_MAJOR)
            t1 = time.perf_counter()
        out.update(nr=nr, nq=nq, d=d, k=k, e2e_ms=(t1 - t0) * 1e3, queries_per_s_e2e=nq / (t1 - t0), stats=st)
        rows = np.random.default_rng(0).choice(nq, 400, replace=False)
        wi, wd = oracle.knn(R, Q[rows], k, method="kdtree")
        out["knn_sample_exact"] = bool((idx[rows] == wi).all() and (dist[rows] == wd).all())
    m.close()
    print(json.dumps(out))
    sys.exit(0)
ctx = sb.Context(0)
if a.config == "D":
    n, d, k = a.n or 4_000_000, 30, 30
    X, _ = gaussian_mixture(n, d, 2)
    ref = torch.from_numpy(X).cuda()
    idx0 = torch.empty((n, k), dtype=torch.int32, device="cuda")
    dist = torch.empty((n, k), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for r in range(a.reps):
        t0 = time.perf_counter()
        s1 = ctx.dev_knn(ref, n, None, n, d, k, idx0, dist)
        t1 = time.perf_counter()
        s2 = ctx.dev_graphs(idx0, n, k, 1 / 15, True, 0, n)
        t2 = time.perf_counter()
    out.update(n=n, d=d, k=k, knn_ms=(t1 - t0) * 1e3, graphs_ms=(t2 - t1) * 1e3,
               cells_per_s=n / (t2 - t0), knn_stats=s1, graph_stats=s2)
    # sampled oracle check of the kNN rows + SNN rows of a block
    rows = np.random.default_rng(0).choice(n, 200, replace=False)
    wi, wd = oracle.knn(X, X[rows], k, method="kdtree")
    gi = idx0[torch.from_numpy(rows).cuda()].cpu().numpy() + 1
    gd = dist[torch.from_numpy(rows).cuda()].cpu().numpy()
    out["knn_sample_exact"] = bool((gi == wi).all() and (gd == wd).all())
    idx1 = (idx0.cpu().numpy() + 1).astype(np.int32)
    p, i, x = oracle.compute_snn_rows(idx1, 1 / 15, 0, 20000)
    r = ctx.dev_result(1)
    P = torch.empty(r.n_rows + 1, dtype=torch.int64, device="cuda")
    I = torch.empty(r.nnz, dtype=torch.int32, device="cuda")
    Xv = torch.empty(r.nnz, dtype=torch.float64, device="cuda")
    ctx.dev_copy_result(1, P, I, Xv)
    Pn = P[:20001].cpu().numpy()
    out["snn_block_exact"] = bool((Pn == p).all() and (I[:p[-1]].cpu().numpy() == i).all()
                                  and (Xv[:p[-1]].cpu().numpy() == x).all())
    out["nnz_per_row"] = r.nnz / n
else:
    nr, nq, d, k = a.n or 2_000_000, (a.n // 4 if a.n else 500_000), 50, 20
    R, c = gaussian_mixture(nr, d, 3)
    Q, _ = gaussian_mixture(nq, d, 4, centres=c)
    ref, qry = torch.from_numpy(R).cuda(), torch.from_numpy(Q).cuda()
    idx0 = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    dist = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for r in range(a.reps):
        t0 = time.perf_counter()
        s1 = ctx.dev_knn(ref, nr, qry, nq, d, k, idx0, dist)
        t1 = time.perf_counter()
    out.update(nr=nr, nq=nq, d=d, k=k, knn_ms=(t1 - t0) * 1e3, queries_per_s=nq / (t1 - t0), knn_stats=s1)
    rows = np.random.default_rng(0).choice(nq, 200, replace=False)
    wi, wd = oracle.knn(R, Q[rows], k, method="kdtree")
    gi = idx0[torch.from_numpy(rows).cuda()].cpu().numpy() + 1
    gd = dist[torch.from_numpy(rows).cuda()].cpu().numpy()
    out["knn_sample_exact"] = bool((gi == wi).all() and (gd == wd).all())
print(json.dumps(out))
