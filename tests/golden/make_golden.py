"""Generates the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference for pbmc_raw.txt):
    python tests/golden/make_golden.py

Everything here is computed WITHOUT the C oracle and WITHOUT the CUDA library:
pure-Python sequential-sum distances (ANN's arithmetic order) and scipy.sparse
for A %*% t(A), so that the oracle and the GPU path are both checked against an
independent restatement of the reference semantics
(/root/reference/src/snn.cpp:14-38, /root/reference/R/clustering.R:664-670).
The n = 6 vectors were derived in SURVEY.md section 8c.
"""
import json
import math
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def snn_scipy(nn1, prune):
    """ComputeSNN restated with scipy: triplets summed, A A^T, v/(k+(k-v)), strict <, drop zeros."""
    nn1 = np.asarray(nn1, dtype=np.int64)
    n, k = nn1.shape
    rows = np.repeat(np.arange(n), k)
    A = sp.coo_matrix((np.ones(n * k), (rows, (nn1 - 1).ravel())), shape=(n, n)).tocsr()  # sums dups
    S = (A @ A.T).tocsc()
    S.sort_indices()
    v = S.data / (k + (k - S.data))
    v[v < prune] = 0.0
    S.data = v
    S.eliminate_zeros()
    S.sort_indices()
    return S.indptr.tolist(), S.indices.tolist(), pack(S.data)


def pack(vals):
    """doubles as a table of distinct C99-hex values + indices (exact and compact)"""
    table = sorted(set(float(v) for v in vals))
    pos = {v: j for j, v in enumerate(table)}
    return {"vals": [v.hex() for v in table], "idx": [pos[float(v)] for v in vals]}


def nn_graph_scipy(nn1, n):
    nn1 = np.asarray(nn1, dtype=np.int64)
    rows_n, k = nn1.shape
    rows = np.repeat(np.arange(rows_n), k)
    A = sp.coo_matrix((np.ones(rows_n * k), (rows, (nn1 - 1).ravel())), shape=(n, n)).tocsc()
    A.sort_indices()
    return A.indptr.tolist(), A.indices.tolist(), pack(A.data)


def knn_python(X, Q, k):
    """Exact kNN, distances summed in dimension order (dist += t*t), ties by lower index."""
    idx, dist = [], []
    for q in Q:
        cand = []
        for r, p in enumerate(X):
            d2 = 0.0
            for a, b in zip(q, p):
                t = a - b
                d2 = d2 + t * t
            cand.append((d2, r))
        cand.sort()
        idx.append([r + 1 for _, r in cand[:k]])
        dist.append([math.sqrt(d2).hex() for d2, _ in cand[:k]])
    return idx, dist


def pbmc_embedding():
    """80 x 10 'pbmc-sized' embedding: pbmc_raw.txt -> log-normalise -> scale -> SVD."""
    path = "/root/reference/inst/extdata/pbmc_raw.txt"
    with open(path) as f:
        lines = f.read().strip().split("\n")
    cells = lines[0].split("\t")
    counts = np.array([[float(v) for v in ln.split("\t")[1:]] for ln in lines[1:]])  # genes x cells
    assert counts.shape == (230, 80) and len(cells) == 80
    lib = counts.sum(axis=0)
    norm = np.log1p(counts / lib * 1e4)
    mu = norm.mean(axis=1, keepdims=True)
    sd = norm.std(axis=1, ddof=1, keepdims=True)
    sd[sd == 0] = 1.0
    scaled = np.clip((norm - mu) / sd, None, 10.0)
    u, s, vt = np.linalg.svd(scaled.T, full_matrices=False)  # cells x genes
    emb = u[:, :10] * s[:10]
    # fix the sign of every component so that the fixture does not depend on LAPACK's choice
    for j in range(10):
        if emb[np.argmax(np.abs(emb[:, j])), j] < 0:
            emb[:, j] = -emb[:, j]
    return cells, emb


def main():
    out = {}
    # ---- SURVEY 8c known-answer vectors (n = 6, k = 3) --------------------
    nn6 = [[1, 2, 3], [2, 1, 3], [3, 2, 4], [4, 5, 3], [5, 4, 3], [6, 5, 4]]
    survey = {
        "nn_ranked": nn6,
        "prune_0": {"p": [0, 5, 10, 16, 22, 28, 32],
                    "i": [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 2, 3, 4, 5],
                    "x": [1, 1, .5, .2, .2, 1, 1, .5, .2, .2, .5, .5, 1, .5, .5, .2, .2, .2, .5, 1, 1, .5,
                          .2, .2, .5, 1, 1, .5, .2, .5, .5, 1]},
        "prune_0.21": {"p": [0, 3, 6, 11, 15, 19, 22],
                       "i": [0, 1, 2, 0, 1, 2, 0, 1, 2, 3, 4, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5],
                       "x": [1, 1, .5, 1, 1, .5, .5, .5, 1, .5, .5, .5, 1, 1, .5, .5, 1, 1, .5, .5, .5, 1]},
    }
    survey["prune_1/15"] = survey["prune_0"]
    # the scipy restatement must reproduce the survey's hand-checked vectors
    for key, prune in (("prune_0", 0.0), ("prune_1/15", 1 / 15), ("prune_0.21", 0.21)):
        p, i, x = snn_scipy(nn6, prune)
        assert p == survey[key]["p"] and i == survey[key]["i"], key
        assert [float.fromhex(x["vals"][j]) for j in x["idx"]] == [float(v) for v in survey[key]["x"]], key
    out["survey_n6"] = survey

    # ---- edge cases pinned with the scipy restatement ----------------------
    rng = np.random.default_rng(12345)
    cases = {}

    def add_case(name, nn1, prune):
        p, i, x = snn_scipy(nn1, prune)
        gp, gi, gx = nn_graph_scipy(nn1, len(nn1))
        cases[name] = {"nn_ranked": np.asarray(nn1).tolist(), "prune": float(prune).hex(),
                       "snn": {"p": p, "i": i, "x": x}, "nn": {"p": gp, "i": gi, "x": gx}}

    def random_lists(n, k, with_self=True):
        nn = np.empty((n, k), dtype=np.int64)
        for r in range(n):
            others = rng.permutation(np.delete(np.arange(n), r))[: k - 1 if with_self else k]
            nn[r] = np.concatenate(([r], others)) if with_self else others
        return nn + 1

    # equality with the cut-off is KEPT (strict <): k=8,s=1 ; k=16,s=2 ; k=24,s=3 all give 1/15
    for k, n in ((8, 60), (16, 120), (24, 150)):
        add_case(f"equality_kept_k{k}", random_lists(n, k), 1 / 15)
    add_case("no_self_k5", random_lists(30, 5, with_self=False), 1 / 15)
    add_case("prune_zero_k10", random_lists(50, 10), 0.0)
    add_case("prune_one_k10", random_lists(50, 10), 1.0)
    add_case("prune_above_one", random_lists(20, 4), 1.5)
    add_case("negative_prune", random_lists(20, 4), -0.5)
    dup = random_lists(25, 6)
    dup[3, 2] = dup[3, 1]          # row 3 lists one neighbour twice  => A = 2 there
    dup[7, 5] = dup[7, 0]          # row 7 lists itself twice
    dup[11, 1:4] = dup[11, 1]      # a triple
    add_case("duplicated_neighbours", dup, 1 / 15)
    add_case("duplicated_neighbours_prune0", dup, 0.0)
    add_case("k_equals_n_minus_1", random_lists(12, 11), 1 / 15)
    add_case("k_equals_1", random_lists(9, 1), 0.0)
    hub = random_lists(60, 6)
    hub[:, 1] = 1                  # everybody lists cell 1: a hub column of length n
    add_case("hub_column", hub, 1 / 15)
    out["snn_cases"] = cases

    # ---- config A: pbmc-sized 80 x 10, k = 20 -------------------------------
    cells, emb = pbmc_embedding()
    idx, dist = knn_python(emb.tolist(), emb.tolist(), 20)
    p, i, x = snn_scipy(idx, 1 / 15)
    gp, gi, gx = nn_graph_scipy(idx, 80)
    out["pbmc80x10"] = {
        "cells": cells,
        "embedding": [[float(v).hex() for v in row] for row in emb],
        "k": 20, "prune": float(1 / 15).hex(),
        "nn_idx": idx, "nn_dist": dist,
        "nn": {"p": gp, "i": gi, "x": gx}, "snn": {"p": p, "i": i, "x": x},
    }
    # query path: first 30 cells as queries against the last 50 as reference
    qidx, qdist = knn_python(emb[30:].tolist(), emb[:30].tolist(), 7)
    out["pbmc_query"] = {"ref_rows": [30, 80], "query_rows": [0, 30], "k": 7, "nn_idx": qidx, "nn_dist": qdist}

    with open(os.path.join(HERE, "fixtures.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "fixtures.json"), os.path.getsize(os.path.join(HERE, "fixtures.json")), "bytes")


if __name__ == "__main__":
    main()
