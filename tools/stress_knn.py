"""kNN stage on differently shaped 1M x 50 inputs: time, visited tiles, uncertified rows, sampled exactness.

    python tools/stress_knn.py [n]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb  # noqa: E402
from seurat_b200.synthetic import shaped_embedding  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d, k = 50, 20


def make(kind):
    return shaped_embedding(kind, n, d)


ctx = sb.Context(0)
idx0 = torch.empty((n, k), dtype=torch.int32, device="cuda")
dist = torch.empty((n, k), dtype=torch.float64, device="cuda")
for kind in sys.argv[2:] or ["mixture32", "manifold10", "pca_like", "dups", "uniform"]:
    X = np.ascontiguousarray(make(kind))
    ref = torch.from_numpy(X).cuda()
    torch.cuda.synchronize()
    best = None
    for r in range(2):
        t0 = time.perf_counter()
        st = ctx.dev_knn(ref, n, None, n, d, k, idx0, dist)
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    rows = torch.randint(0, n, (128,), device="cuda")
    d2 = torch.cdist(ref[rows], ref, compute_mode="donot_use_mm_for_euclid_dist") ** 2
    want = d2.topk(k, largest=False)
    got = idx0[rows].long()
    got_d = dist[rows]
    # compare distances (ties may legitimately reorder equal-distance neighbours between the two methods)
    ok = torch.allclose(want.values.sqrt(), got_d, rtol=1e-9, atol=1e-9)
    print(f"{kind:11s} {best:8.1f} ms  prep {st['ms_prep']:.1f} cand {st['ms_knn_cand']:.1f} rerank {st['ms_knn_rerank']:.1f} "
          f"fallback {st['ms_knn_fallback']:.1f}  uncertified {st['n_uncertified']}  tiles {st['knn_tiles_visited']}/"
          f"{st['knn_tiles_total']}  sampled distances match: {ok}", flush=True)
