"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb  # noqa: E402
from seurat_b200 import _lib  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
ctx = sb.Context(0)
X, c = gaussian_mixture(n, 50, seed=5, n_centres=8)
for mode in (1, 2):
    idx, dist, st = ctx.knn(X, None, 20, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, mode)
    print("mode", mode, "uncertified", st["n_uncertified"], "tiles", st["knn_tiles_visited"], flush=True)
# many fallback rows (grouped FP64 scan) and a query != reference call
idx2, dist2, st = ctx.knn(X, None, 20, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 21, 2)
print("narrow candidates: uncertified", st["n_uncertified"], "same result:", bool((idx2 == idx).all()), flush=True)
Q, _ = gaussian_mixture(777, 50, seed=6, centres=c)
ctx.knn(X, Q, 20, _lib.ROW_MAJOR, _lib.KNN_TENSOR, 0, 2)
# full path: kNN + nn Graph + SNN (filtered kernels, all row classes), then plain kernels (prune 0)
r = ctx.find_neighbors(X, 20, 1 / 15, True, _lib.ROW_MAJOR)
print("snn nnz", r["snn"][1].size, flush=True)
p, i, x, st = ctx.compute_snn(r["idx"][:8000] % 8000 + 1 if False else (idx[:8000] - 1) % 8000 + 1, 0.0, _lib.ROW_MAJOR)
print("prune 0 nnz", i.size, flush=True)
ix = ctx.index(X, _lib.ROW_MAJOR)
ix.knn(Q, 5)
ix.close()
print("done")
