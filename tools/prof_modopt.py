"""FindNeighbors + FindClusters on the device-resident SNN graph: time and statistics of the modularity
optimiser (b200nn_modularity_cluster_resident) at a given size, optionally beside the real reference
(oracle/_ref) on the same graph.

    python tools/prof_modopt.py [--n 1000000] [--starts 1] [--iters 10] [--algorithm 1] [--ref]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb  # noqa: E402
from seurat_b200 import _lib  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=int, default=50)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--starts", type=int, default=1)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--algorithm", type=int, default=1)
ap.add_argument("--resolution", type=float, default=0.8)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--ref", action="store_true", help="also run the real reference (oracle/_ref) on the same graph")
a = ap.parse_args()

X, _ = gaussian_mixture(a.n, a.d, a.seed)
ctx = sb.Context(0)
t0 = time.perf_counter()
r = ctx.find_neighbors(X, a.k, 1 / 15, True, _lib.ROW_MAJOR)
print(f"find_neighbors: {time.perf_counter() - t0:.2f} s, nnz_snn {r['stats']['nnz_snn']}", flush=True)
lab = None
for rep in range(a.reps):
    t0 = time.perf_counter()
    lab, st = ctx.modularity_cluster_resident(a.n, 1, a.resolution, a.algorithm, a.starts, a.iters, 0)
    print(json.dumps({"rep": rep, "wall_s": time.perf_counter() - t0, **st}), flush=True)
if a.ref:
    import oracle
    p, i, x = r["snn"]
    t0 = time.perf_counter()
    want, nc, q = oracle.modularity_cluster_ref(p, i, x, 1, a.resolution, a.algorithm, a.starts, a.iters, 0)
    print(json.dumps({"reference_s": time.perf_counter() - t0, "n_clusters": nc, "modularity": q,
                      "identical_labels": bool((want == lab).all()), "identical_modularity": q == st["modularity"]}),
          flush=True)
ctx.close()
