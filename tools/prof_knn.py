"""Profiling driver for the tensor-core candidate kernel (K1) alone.

    python tools/prof_knn.py --nr 1000000 --nq 37888 --d 50 --kp 32 --reps 3

Runs b200nn_dev_knn_candidates (prep + tcgen05 kernel, no re-rank) on synthetic
mixture data; nq = 148 * 256 gives every SM exactly one work item.  Used under
`ncu -k regex:k_knn_tc` to keep the profiled launch short.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nr", type=int, default=1_000_000)
ap.add_argument("--nq", type=int, default=148 * 256)
ap.add_argument("--d", type=int, default=50)
ap.add_argument("--kp", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 exhaustive, 2 grouped")
ap.add_argument("--sort", default="none", choices=["none", "cluster", "pc1"])
a = ap.parse_args()

X, centres = gaussian_mixture(a.nr, a.d, a.seed)
if a.sort == "cluster":   # perfect spatial coherence: points grouped by their mixture component
    lab = np.empty(a.nr, dtype=np.int64)
    for s0 in range(0, a.nr, 100000):
        blk = X[s0:s0 + 100000]
        d2 = (blk ** 2).sum(1)[:, None] - 2 * blk @ centres.T + (centres ** 2).sum(1)[None]
        lab[s0:s0 + 100000] = d2.argmin(1)
    X = np.ascontiguousarray(X[np.argsort(lab, kind="stable")])
elif a.sort == "pc1":
    X = np.ascontiguousarray(X[np.argsort(X[:, 0], kind="stable")])
ctx = sb.Context(0)
ref = torch.from_numpy(X).cuda()
if a.nq < a.nr:   # a contiguous block of queries from the middle (coherent when sorted)
    o = (a.nr - a.nq) // 2
    qry = ref[o:o + a.nq].clone()
else:
    qry = ref
cand = torch.empty((a.nq, a.kp), dtype=torch.int32, device="cuda")
tau = torch.empty(a.nq, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for r in range(a.reps):
    t0 = time.perf_counter()
    ctx.dev_knn_candidates(ref, a.nr, qry, a.nq, a.d, a.kp, cand, tau, None, None, a.mode)
    dt = time.perf_counter() - t0
    print(f"rep {r}: {dt * 1e3:.2f} ms  ({a.nq * a.nr / dt / 1e12:.3f} T candidates/s, "
          f"{2 * a.nq * a.nr * a.d / dt / 1e12:.1f} algorithmic TFLOP/s)")
