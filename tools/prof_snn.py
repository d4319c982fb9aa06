"""Profiling driver for the nn-Graph + SNN stage (K3/K4) alone.

    python tools/prof_snn.py --n 1000000 --d 50 --k 20 --reps 3

Builds the kNN index on the GPU once (untimed), then runs b200nn_dev_graphs.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=int, default=50)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--prune", type=float, default=1 / 15)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()

X, _ = gaussian_mixture(a.n, a.d, a.seed)
ctx = sb.Context(0)
ref = torch.from_numpy(X).cuda()
idx0 = torch.empty((a.n, a.k), dtype=torch.int32, device="cuda")
dist = torch.empty((a.n, a.k), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ctx.dev_knn(ref, a.n, None, a.n, a.d, a.k, idx0, dist)
for r in range(a.reps):
    t0 = time.perf_counter()
    st = ctx.dev_graphs(idx0, a.n, a.k, a.prune, True, 0, a.n)
    dt = time.perf_counter() - t0
    print(f"rep {r}: {dt * 1e3:.2f} ms  nn_graph {st['ms_nn_graph']:.2f} ms  snn {st['ms_snn']:.2f} ms  "
          f"nnz {st['nnz_snn']}  expanded {st['snn_expanded']}  max indegree {st['max_indegree']}")
