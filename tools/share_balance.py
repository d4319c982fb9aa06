"""Balance of the group-aligned kNN shares (b200nn_opts.shard_rank / shard_world) on ONE device: per share the
number of rows, the (item, tile) pairs visited and the stage times.   python tools/share_balance.py [world ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

n, d, k = 1_000_000, 50, 20
X, _ = gaussian_mixture(n, d, 1)
ctx = sb.Context(0)
ref = torch.from_numpy(X).cuda()
idx = torch.zeros((n, k), dtype=torch.int32, device="cuda")
dd = torch.empty((n, k), dtype=torch.float64, device="cuda")
for world in [int(v) for v in sys.argv[1:]] or [4, 8]:
    for rank in range(world):
        for rep in range(2):
            dd.fill_(float("nan"))
            st = ctx.dev_knn(ref, n, None, n, d, k, idx, dd, shard_rank=rank, shard_world=world)
        rows = int((~torch.isnan(dd[:, 0])).sum())
        print(f"world {world} rank {rank}: rows {rows}  tiles {st['knn_tiles_visited']}  prep {st['ms_prep']:.2f}  "
              f"cand {st['ms_knn_cand']:.2f} (pass 2 {st['ms_knn_pass2']:.2f})  rerank {st['ms_knn_rerank']:.2f}  "
              f"fallback {st['ms_knn_fallback']:.2f}  uncertified {st['n_uncertified']}", flush=True)
ctx.close()
