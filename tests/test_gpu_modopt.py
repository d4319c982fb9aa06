"""GPU suite (-m gpu): the modularity optimiser on the device, through the C ABI.

Parity bar: the cluster vector AND the bits of the modularity equal the reference's for the same
seed.  Checked against (i) the known answers the reference's own tests hold, (ii) the committed
golden vectors the real reference produced (tests/golden/make_modopt_golden.py), (iii) the real
reference itself (oracle/_ref/libmodopt_ref.so travels to the box as a prebuilt file) on random
graphs, and (iv) on the SNN graph FindNeighbors left in HBM (config B, 100k cells).
"""
import json
import os

import numpy as np
import pytest

import modopt_cases as mc
import oracle
import seurat_b200 as sb
from seurat_b200 import _lib
from seurat_b200.synthetic import CONFIGS, gaussian_mixture

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "modopt_fixtures.json")) as f:
    GOLD = json.load(f)["cases"]
GRAPHS = {name: build for name, build, _ in mc.small_cases()}
PRUNE = 1 / 15


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def test_reference_tests_known_answers(ctx):
    """tests/testthat/test_modularity_optimizer.R: 'Algorithm 1', 'Algorithm 2', 'Algorithm 3' and 'Low
    Resolution' on the karate club"""
    p, i, x = mc.slots(mc.karate())
    for params, want in mc.KARATE_KAT + mc.KARATE_KAT_SLM:
        got, st = ctx.modularity_cluster(p, i, x, *params)
        assert got.tolist() == want and st["n_clusters"] == max(want) + 1


@pytest.mark.parametrize("case", [c for c in GOLD if not c["graph"].startswith("snn:")],
                         ids=lambda c: c["graph"] + "-" + "-".join(map(str, c["params"])))
def test_matches_golden(ctx, case, monkeypatch):
    p, i, x = mc.slots(GRAPHS[case["graph"]]())
    for w0, wm in ((None, None), ("1", "1"), ("3", "64")):
        if w0 is None:
            monkeypatch.delenv("B200NN_MODOPT_WIN0", raising=False)
            monkeypatch.delenv("B200NN_MODOPT_WINMAX", raising=False)
        else:
            if p.size - 1 > 400:
                continue                     # node-by-node execution: small graphs only
            monkeypatch.setenv("B200NN_MODOPT_WIN0", w0)
            monkeypatch.setenv("B200NN_MODOPT_WINMAX", wm)
        lab, st = ctx.modularity_cluster(p, i, x, *case["params"])
        assert st["n_clusters"] == case["n_clusters"], st
        assert mc.digest(lab) == case["sha256"], st
        assert st["modularity"] == float.fromhex(case["modularity"]), "modularity must agree to the last bit"
        assert st["gpu_launches"] > 0


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libmodopt_ref.so missing")
@pytest.mark.parametrize("seed", range(4))
def test_matches_the_reference_on_random_graphs(ctx, seed):
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.integers(200, 6000))
    M = mc.random_graph(n, float(rng.uniform(3, 20)), seed=seed + 80, weighted=bool(seed % 2),
                        symmetric=bool(seed % 3), isolated=int(rng.integers(0, 4)))
    p, i, x = mc.slots(M)
    for _ in range(3):
        mf = int(rng.integers(1, 3))
        res = float(rng.choice([0.01, 0.2, 1.0])) if mf == 2 else float(rng.choice([0.1, 0.8, 1.0, 4.0, 40.0]))
        params = (mf, res, int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(1, 6)),
                  int(rng.integers(0, 2 ** 40)))
        want, wc, wq = oracle.modularity_cluster_ref(p, i, x, *params)
        got, st = ctx.modularity_cluster(p, i, x, *params)
        assert st["n_clusters"] == wc and (got == want).all() and st["modularity"] == wq, (params, st)


def test_snn_20k_golden_and_python_mirror(ctx):
    X, _ = gaussian_mixture(20000, 50, seed=7, n_centres=12)
    r = ctx.find_neighbors(X, 20, PRUNE, True, _lib.ROW_MAJOR)
    p, i, x = r["snn"]
    for c in [c for c in GOLD if c["graph"] == "snn:20000:50:7:12:20"]:
        lab, st = ctx.modularity_cluster(p, i, x, *c["params"])
        assert st["n_clusters"] == c["n_clusters"] and mc.digest(lab) == c["sha256"]
        assert st["modularity"] == float.fromhex(c["modularity"])
    c = [c for c in GOLD if c["graph"] == "snn:20000:50:7:12:20"][0]
    mf, res, alg, ns, ni, seed = c["params"]
    g = sb.Graph(p, i, x, (20000, 20000))
    ids = sb.RunModularityClustering(g, mf, res, alg, ns, ni, seed, print_output=False, ctx=ctx)
    assert mc.digest(ids) == c["sha256"]
    out = sb.FindClusters(g, modularity_fxn=mf, resolution=[res], algorithm="louvain", n_start=ns, n_iter=ni,
                          random_seed=seed, verbose=False, ctx=ctx)
    assert mc.digest(out[f"res.{res}"]) == c["sha256"]          # no singletons in this graph


def test_resident_snn_config_b(ctx):
    """FindNeighbors on config B (100k x 50) leaves the SNN graph in HBM; clustering it in place gives what
    the reference gives on the same graph (golden made from the oracle's SNN, which the device SNN equals)"""
    cfg = CONFIGS["B"]
    X, _ = gaussian_mixture(cfg["n"], cfg["d"], cfg["seed"])
    ctx.find_neighbors(X, cfg["k"], PRUNE, True, _lib.ROW_MAJOR)
    for c in [c for c in GOLD if c["graph"] == "snn:100000:50:0:32:20"]:      # Louvain and SLM
        lab, st = ctx.modularity_cluster_resident(cfg["n"], *c["params"])
        assert st["n_clusters"] == c["n_clusters"] and mc.digest(lab) == c["sha256"], st
        assert st["modularity"] == float.fromhex(c["modularity"])
        assert st["rounds"] <= 12 * st["windows"] + 50, st
    with pytest.raises(sb.B200Error, match="expected"):
        ctx.modularity_cluster_resident(cfg["n"] + 1, *c["params"])


def test_argument_errors_carry_the_wrappers_messages(ctx):
    p, i, x = mc.slots(mc.karate())
    for kw, msg in (({"modularity_function": 3}, "Modularity parameter must be equal to 1 or 2"),
                    ({"algorithm": 5}, "Algorithm for modularity optimization must be 1, 2, 3, or 4"),
                    ({"n_random_starts": 0}, "Have to have at least one start"),
                    ({"n_iterations": 0}, "Need at least one interation"),
                    ({"modularity_function": 2, "resolution": 1.5}, "resolution<1 for alternative modularity"),
                    ({"algorithm": 4}, "Leiden")):
        with pytest.raises(sb.B200Error, match=msg):
            ctx.modularity_cluster(p, i, x, **kw)
    with pytest.raises(sb.B200Error, match="Matrix contained no network data"):
        ctx.modularity_cluster(np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0))
    lab, _ = ctx.modularity_cluster(p, i, x, 1, 1.0, 1, 1, 1, 564)      # the context keeps working
    assert lab.tolist() == mc.KARATE_KAT[0][1]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libmodopt_ref.so missing")
def test_edge_file_input_matches_the_reference_reading_the_same_file(ctx, tmp_path):
    """edge.file.name of RunModularityClusteringCpp (RModularityOptimizer.cpp:55-63) through
    b200nn_modularity_cluster_edge_file / _edges and through the host mirror"""
    rng = np.random.default_rng(5)
    n, m = 3000, 20000
    n1, n2 = rng.integers(0, n, m), rng.integers(0, n, m)
    w = np.round(rng.random(m) + 0.01, 6)
    path = str(tmp_path / "edges.txt")
    with open(path, "w") as f:
        for e in range(m):
            f.write(f"{n1[e]}\t{n2[e]}\t{float(w[e])!r}\n")
    for alg in (1, 2, 3):
        params = (1, 0.8, alg, 2, 5, 42)
        want, wc, wq = oracle.modularity_cluster_file_ref(path, *params)
        got, st = ctx.modularity_cluster_edge_file(path, *params)
        assert (got == want).all() and st["n_clusters"] == wc and st["modularity"] == wq
        got2, st2 = ctx.modularity_cluster_edges(n1, n2, w, want.size, *params)
        assert (got2 == want).all() and st2["modularity"] == wq
    ids = sb.RunModularityClustering(None, 1, 0.8, 1, 2, 5, 42, print_output=False, edge_file_name=path, ctx=ctx)
    want, _, _ = oracle.modularity_cluster_file_ref(path, 1, 0.8, 1, 2, 5, 42)
    assert (ids == want).all()
    # unweighted file (weight 1), and the failures the wrapper reports
    path2 = str(tmp_path / "unweighted.txt")
    with open(path2, "w") as f:
        for e in range(m):
            f.write(f"{n1[e]}\t{n2[e]}\n")
    want, wc, wq = oracle.modularity_cluster_file_ref(path2, 1, 0.8, 1, 1, 3, 7)
    got, st = ctx.modularity_cluster_edge_file(path2, 1, 0.8, 1, 1, 3, 7)
    assert (got == want).all() and st["modularity"] == wq
    with pytest.raises(sb.B200Error, match="Could not parse edge file"):
        ctx.modularity_cluster_edge_file(str(tmp_path / "missing.txt"))
    bad = str(tmp_path / "bad.txt")
    with open(bad, "w") as f:
        f.write("0\t1\t0.5\nx\t2\t0.1\n")
    with pytest.raises(sb.B200Error, match="Could not parse edge file"):
        ctx.modularity_cluster_edge_file(bad)
