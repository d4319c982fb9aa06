"""Config C through b200nn_multi_* with the shards sharing ONE device (emulation of the multi-GPU path on a
single-GPU box): compares with the single-device result and prints the phase trace (B200NN_TRACE=1).

    python tools/multi_emul.py [n_shards] [n]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb  # noqa: E402
from seurat_b200 import _lib  # noqa: E402
from seurat_b200.synthetic import gaussian_mixture  # noqa: E402

shards = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
X, _ = gaussian_mixture(n, 50, 1)
single = sb.Context(0)
t0 = time.perf_counter()
want = single.find_neighbors(X, 20, 1 / 15, True, _lib.ROW_MAJOR)
print(f"single device: {time.perf_counter() - t0:.2f} s", flush=True)
m = sb.MultiContext([0] * shards)
for rep in range(2):
    t0 = time.perf_counter()
    got = m.find_neighbors(X, 20, 1 / 15, True, _lib.ROW_MAJOR)
    print(f"{shards} shards on device 0: {time.perf_counter() - t0:.2f} s  stats {got['stats']}", flush=True)
ok = bool((got["idx"] == want["idx"]).all() and (got["dist"] == want["dist"]).all())
for a, b in zip(got["snn"] + got["nn"], want["snn"] + want["nn"]):
    ok = ok and a.shape == b.shape and bool((a == b).all())
print("identical to the single-device result:", ok, flush=True)
m.close()
single.close()
