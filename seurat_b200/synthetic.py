"""Synthetic 'PCA-like' embeddings used by the tests and bench.py.

Gaussian mixture: C = 32 centres ~ N(0, 6^2 I_d) drawn first, then labels,
then unit-variance noise; float64, one cell per row, names cell{i}.
"""
from __future__ import annotations

import numpy as np

# config name -> (n, d, k, seed) as fixed in SURVEY.md 8d
CONFIGS = {
    "B": dict(n=100_000, d=50, k=20, seed=0),
    "C": dict(n=1_000_000, d=50, k=20, seed=1),
    "D": dict(n=4_000_000, d=30, k=30, seed=2),
}


def gaussian_mixture(n: int, d: int, seed: int, n_centres: int = 32, centre_sd: float = 6.0,
                     centres: np.ndarray | None = None, chunk: int = 250_000):
    """Returns (X float64 [n, d] C-order, centres).  Generated in chunks to bound memory."""
    rng = np.random.default_rng(seed)
    if centres is None:
        centres = rng.normal(0.0, centre_sd, size=(n_centres, d))
    lab = rng.integers(0, centres.shape[0], size=n)
    X = np.empty((n, d), dtype=np.float64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        X[s:e] = centres[lab[s:e]] + rng.normal(0.0, 1.0, size=(e - s, d))
    return X, centres


def cell_names(n: int) -> list[str]:
    return [f"cell{i}" for i in range(n)]
