"""kNN stage at k > 32 (column splits of the grouped search) on the benchmark input: time, uncertified rows,
sampled rows against the oracle.   python tools/prof_large_k.py [--n 1000000] [--d 50] [--ks 50,100,200]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker only)
import seurat_b200 as sb  # noqa: E402
from seurat_b200 import _lib  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=int, default=50)
ap.add_argument("--ks", default="50,100,200")
ap.add_argument("--check", type=int, default=500)
a = ap.parse_args()
import torch  # noqa: E402

X, _ = gaussian_mixture(a.n, a.d, 1)
ctx = sb.Context(0)
ref = torch.from_numpy(X).cuda()
for k in [int(v) for v in a.ks.split(",")]:
    idx = torch.empty((a.n, k), dtype=torch.int32, device="cuda")
    dd = torch.empty((a.n, k), dtype=torch.float64, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = ctx.dev_knn(ref, a.n, None, a.n, a.d, k, idx, dd)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
    rows = np.sort(np.random.default_rng(k).choice(a.n, a.check, replace=False))
    wi, wd = oracle.knn(X, X[rows], k, method="kdtree")
    ok = bool((idx[rows].cpu().numpy() + 1 == wi).all() and (dd[rows].cpu().numpy() == wd).all())
    print(f"k={k}: {ms:.1f} ms  prep {st['ms_prep']:.1f} cand {st['ms_knn_cand']:.1f} rerank {st['ms_knn_rerank']:.1f} "
          f"fallback {st['ms_knn_fallback']:.1f}  uncertified {st['n_uncertified']}  tiles {st['knn_tiles_visited']}"
          f"/{st['knn_tiles_total']}  {a.check} sampled rows exact: {ok}", flush=True)
ctx.close()
